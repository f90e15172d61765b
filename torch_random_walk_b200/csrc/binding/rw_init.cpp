// rw_init.cpp -- the pybind11 module `torch_rw_native`, B200 edition.
//
// Successor of the reference's csrc/rw_init.cpp (module definition at :133-141): the same seven
// functions with the same positional signatures, but no device dispatch and no kernels of its own.
// Each function does what the reference's launchers did around their kernels -- CHECK_CUDA
// (csrc/cuda/utils.cuh:7-9), torch::empty on the inputs' device, current-stream lookup -- and then
// calls the C ABI of libtrw_b200.so (include/trw_b200.h).  CPU tensors raise: there is no CPU path.
//
// Optional: the shipped Python binding (torch_random_walk_b200/native.py, ctypes) needs no compiler
// on the user's machine; this file is the drop-in for a maintainer who keeps the reference's
// extension layout.  Built by torch_random_walk_b200/_build_ext.py, tested in tests/test_ext.py.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include <tuple>

#include "trw_b200.h"

#define CHECK_CUDA(x) TORCH_CHECK((x).is_cuda(), #x " must be a CUDA tensor")
#define CHECK_CONTIGUOUS(x) TORCH_CHECK((x)->is_contiguous(), #x " must be a contigous tensor")
#define TRW_CHECK(call) TORCH_CHECK((call) == TRW_OK, trw_last_error())

namespace {

const int64_t* ptr(const torch::Tensor& t) { return t.numel() ? t.data_ptr<int64_t>() : nullptr; }
void* stream() { return at::cuda::getCurrentCUDAStream().stream(); }
torch::TensorOptions like(const torch::Tensor& t) { return torch::TensorOptions().dtype(torch::kInt64).device(t.device()); }

}  // namespace

torch::Tensor walk(const torch::Tensor* row_ptr, const torch::Tensor* column_idx, const torch::Tensor* target_nodes,
                   const double p, const double q, const int walk_length, const int seed) {
  CHECK_CUDA((*row_ptr));
  CHECK_CUDA((*column_idx));
  CHECK_CUDA((*target_nodes));
  c10::cuda::CUDAGuard guard(row_ptr->device());
  auto rp = row_ptr->contiguous(), ci = column_idx->contiguous(), tg = target_nodes->contiguous();
  auto walks = torch::empty({tg.size(0), walk_length + 1}, like(rp));
  const int64_t n_nodes = std::max<int64_t>(rp.size(0) - 1, 0), nnz = ci.size(0);
  const size_t need = tg.size(0) ? trw_walk_csr_workspace_bytes_for(n_nodes, nnz, p, q, tg.size(0), (int)walk_length) : 0;
  auto ws = torch::empty({(int64_t)need}, like(rp).dtype(torch::kUInt8));
  TRW_CHECK(trw_walk_csr(ptr(rp), ptr(ci), n_nodes, nnz, ptr(tg), tg.size(0), 0, p, q, walk_length, seed,
                         walks.data_ptr<int64_t>(), walk_length + 1, need ? ws.data_ptr() : nullptr, need,
                         rp.device().index(), stream()));
  return walks;
}

torch::Tensor walk_edge_list(const torch::Tensor* edge_list_indexed, const torch::Tensor* node_edges_idx,
                             const torch::Tensor* target_nodes, const double p, const double q, const int walk_length,
                             const int seed, const int64_t padding_idx, const bool restart) {
  CHECK_CUDA((*edge_list_indexed));
  CHECK_CUDA((*node_edges_idx));
  CHECK_CUDA((*target_nodes));
  c10::cuda::CUDAGuard guard(node_edges_idx->device());
  auto el = edge_list_indexed->contiguous(), nei = node_edges_idx->contiguous(), tg = target_nodes->contiguous();
  auto walks = torch::empty({tg.size(0), walk_length + 1}, like(nei));
  const size_t need = tg.size(0) ? trw_walk_edge_list_workspace_bytes(el.size(0), nei.size(0), p, q) : 0;
  auto ws = torch::empty({(int64_t)need}, like(nei).dtype(torch::kUInt8));
  TRW_CHECK(trw_walk_edge_list_ws(ptr(el), el.size(0), ptr(nei), nei.size(0), ptr(tg), tg.size(0), 0, p, q, walk_length, seed,
                                  padding_idx, restart ? 1 : 0, walks.data_ptr<int64_t>(), walk_length + 1,
                                  need ? ws.data_ptr() : nullptr, need, nei.device().index(), stream()));
  return walks;
}

torch::Tensor walk_triples(const torch::Tensor* triples_indexed, const torch::Tensor* relation_tail_index,
                           const torch::Tensor* target_nodes, const int walk_length, const int64_t padding_idx,
                           const bool restart, const int seed) {
  CHECK_CUDA((*triples_indexed));
  CHECK_CUDA((*relation_tail_index));
  CHECK_CUDA((*target_nodes));
  c10::cuda::CUDAGuard guard(target_nodes->device());
  auto tr = triples_indexed->contiguous(), rti = relation_tail_index->contiguous(), tg = target_nodes->contiguous();
  auto walks = torch::empty({tg.size(0), 2 * walk_length + 1}, like(tg));
  TRW_CHECK(trw_walk_triples(ptr(tr), tr.size(0), ptr(rti), rti.size(0), ptr(tg), tg.size(0), 0, walk_length, padding_idx,
                             restart ? 1 : 0, seed, walks.data_ptr<int64_t>(), 2 * walk_length + 1, tg.device().index(),
                             stream()));
  return walks;
}

template <typename Fn>
static std::tuple<at::Tensor, at::Tensor, at::Tensor> node_windows(Fn fn, const torch::Tensor* walks, int window_size,
                                                                   int64_t num_nodes, int seed, bool cbow) {
  CHECK_CUDA((*walks));
  CHECK_CONTIGUOUS(walks);
  c10::cuda::CUDAGuard guard(walks->device());
  const int64_t n = walks->size(0), wl = walks->size(1), k = (wl - window_size + 1) * n;
  auto first = torch::empty({k}, like(*walks));
  auto win = torch::empty({k, window_size - 1}, like(*walks));
  auto other = cbow ? torch::empty({k}, like(*walks)) : torch::empty({k, window_size - 1}, like(*walks));
  auto& o1 = cbow ? other : win;
  auto& o2 = cbow ? win : other;
  TRW_CHECK(fn(ptr(*walks), n, wl, window_size, num_nodes, seed, first.data_ptr<int64_t>(), o1.data_ptr<int64_t>(),
               o2.data_ptr<int64_t>(), walks->device().index(), stream()));
  return std::make_tuple(first, o1, o2);
}

std::tuple<at::Tensor, at::Tensor, at::Tensor> to_windows(const torch::Tensor* walks, const int window_size,
                                                          const int64_t num_nodes, const int seed) {
  return node_windows(trw_windows, walks, window_size, num_nodes, seed, false);
}

std::tuple<at::Tensor, at::Tensor, at::Tensor> to_windows_cbow(const torch::Tensor* walks, const int window_size,
                                                               const int64_t num_nodes, const int seed) {
  return node_windows(trw_windows_cbow, walks, window_size, num_nodes, seed, true);
}

template <typename Fn>
static std::tuple<at::Tensor, at::Tensor, at::Tensor> triple_windows(Fn fn, const torch::Tensor* walks, int window_size,
                                                                     int64_t num_nodes, int64_t padding_idx,
                                                                     const torch::Tensor* triples, int seed, bool cbow) {
  CHECK_CUDA((*walks));
  CHECK_CONTIGUOUS(walks);
  CHECK_CUDA((*triples));
  c10::cuda::CUDAGuard guard(walks->device());
  auto tr = triples->contiguous();
  const int64_t n = walks->size(0), wl = walks->size(1), k = ((wl - 1) / 2) * n;
  auto first = torch::empty({k, 3}, like(*walks));
  auto win = torch::empty({k, 2 * window_size, 3}, like(*walks));
  auto other = cbow ? torch::empty({k, 3}, like(*walks)) : torch::empty({k, 2 * window_size, 3}, like(*walks));
  auto& o1 = cbow ? other : win;
  auto& o2 = cbow ? win : other;
  // workspace for the 16-byte copy of `triples` the negative rows are gathered from
  const size_t need = trw_windows_triples_workspace_bytes(tr.size(0));
  auto ws = torch::empty({(int64_t)need}, like(*walks).dtype(torch::kUInt8));
  TRW_CHECK(fn(ptr(*walks), n, wl, window_size, num_nodes, padding_idx, ptr(tr), tr.size(0), seed, first.data_ptr<int64_t>(),
               o1.data_ptr<int64_t>(), o2.data_ptr<int64_t>(), need ? ws.data_ptr() : nullptr, need, walks->device().index(),
               stream()));
  return std::make_tuple(first, o1, o2);
}

std::tuple<at::Tensor, at::Tensor, at::Tensor> to_windows_triples(const torch::Tensor* walks, const int window_size,
                                                                  const int64_t num_nodes, const int64_t padding_idx,
                                                                  const torch::Tensor* triples, const int seed) {
  return triple_windows(trw_windows_triples_ws, walks, window_size, num_nodes, padding_idx, triples, seed, false);
}

std::tuple<at::Tensor, at::Tensor, at::Tensor> to_windows_triples_cbow(const torch::Tensor* walks, const int window_size,
                                                                       const int64_t num_nodes, const int64_t padding_idx,
                                                                       const torch::Tensor* triples, const int seed) {
  return triple_windows(trw_windows_triples_cbow_ws, walks, window_size, num_nodes, padding_idx, triples, seed, true);
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("walk", &walk, "walk");
  m.def("walk_edge_list", &walk_edge_list, "walk_edge_list");
  m.def("walk_triples", &walk_triples, "walk_triples");
  m.def("to_windows", &to_windows, "to_windows");
  m.def("to_windows_cbow", &to_windows_cbow, "to_windows");
  m.def("to_windows_triples", &to_windows_triples, "to_windows_triples");
  m.def("to_windows_triples_cbow", &to_windows_triples_cbow, "to_windows_triples_cbow");
}
