/*
 * trw_oracle.c -- CPU restatement of torch_rw's walk sampler and window generator.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under torch_random_walk_b200/ may include, link or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker or as the timed CPU baseline.
 *
 * Parity status: PINNED.  Every function below draws from glibc srand()/rand() in exactly the
 * order the reference's csrc/cpu implementation does, so (single-threaded, like the reference
 * on small inputs) it reproduces the reference's own CPU golden vectors bit for bit:
 *   tests/test_rw.py:30-55,98-122, tests/test_rw_edge_list.py (8 CPU tests),
 *   tests/test_rw_triples.py:12-81, tests/test_windows.py:4-31,34-55,122-180,243-285
 * (checked in tests/test_oracle_golden.py), and outputs of the unmodified reference built into
 * oracle/_ref on random inputs (tests/golden/ref_*.npz, made by tests/golden/make_golden.py).
 *
 * All arrays are int64, row-major, contiguous.  Citations are reference file:line.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

typedef int64_t i64;

/* csrc/cpu/cpu_utils.cpp:3-10 -- inclusive range; NO draw when the range has one value. */
static i64 sample_int(i64 start, i64 end)
{
    if (start == end) return start;
    return start + ((i64)rand() % ((end + 1) - start));
}

/* csrc/cpu/rw_cpu.cpp:119-123 (same in rw_cpu_edge_list.cpp:164-168). */
static void rejection_probs(double p, double q, double *p0, double *p1, double *p2)
{
    double max_prob_init = fmax(1.0 / p, 1);
    double max_prob = fmax(max_prob_init, 1.0 / q);
    *p0 = 1.0 / p / max_prob;
    *p1 = 1.0 / max_prob;
    *p2 = 1.0 / q / max_prob;
}

/* ------------------------------------------------------------------ CSR walk */

/* csrc/cpu/rw_cpu.cpp:7-30.  One rand() per call.  The CPU source divides by zero on a
 * degree-0 node (SIGFPE); here the draw is still consumed and the node is kept, which is the
 * explicit out-of-range branch of the source (rw_cpu.cpp:27-29, csrc/cuda/rw_cuda.cu:25-30). */
static i64 csr_sample_neighbor(i64 node, const i64 *row_ptr, const i64 *col_idx, i64 nnz)
{
    i64 column_start = row_ptr[node];
    i64 column_end = row_ptr[node + 1];
    i64 r = (i64)rand();
    i64 deg = column_end - column_start;
    if (deg == 0) return node;
    i64 nbr_idx = column_start + (r % deg);
    if (nbr_idx >= 0 && nbr_idx < nnz) return col_idx[nbr_idx];
    return node;
}

/* csrc/cpu/rw_cpu.cpp:32-56 -- linear scan of adj(previous_node). */
static int csr_is_neighbor(i64 new_node, i64 previous_node, const i64 *row_ptr, const i64 *col_idx)
{
    for (i64 i = row_ptr[previous_node]; i < row_ptr[previous_node + 1]; i++)
        if (col_idx[i] == new_node) return 1;
    return 0;
}

/* walk_cpu, csrc/cpu/rw_cpu.cpp:203-226; uniform_walk :58-107; biased_walk :109-201.
 * out is [n_walks, walk_length+1]. */
int orc_walk_csr(const i64 *row_ptr, const i64 *col_idx, i64 nnz, const i64 *targets, i64 n_walks,
                 double p, double q, int walk_length, int seed, i64 *out)
{
    i64 ws = (i64)walk_length + 1;
    srand((unsigned)seed);
    if (p == 1.0 && q == 1.0) {
        for (i64 w = 0; w < n_walks; w++) {
            i64 *row = out + w * ws;
            i64 prev = targets[w];
            row[0] = prev;
            for (i64 s = 1; s < ws; s++) {
                prev = csr_sample_neighbor(prev, row_ptr, col_idx, nnz);
                row[s] = prev;
            }
        }
        return 0;
    }
    double prob_0, prob_1, prob_2;
    rejection_probs(p, q, &prob_0, &prob_1, &prob_2);
    for (i64 w = 0; w < n_walks; w++) {
        i64 *row = out + w * ws;
        row[0] = targets[w];
        if (ws < 2) continue; /* the source writes row[1] out of bounds here (rw_cpu.cpp:159) */
        row[1] = csr_sample_neighbor(targets[w], row_ptr, col_idx, nnz);
        i64 prev = row[1];
        for (i64 s = 2; s < ws; s++) {
            i64 selected;
            for (;;) {
                i64 x = csr_sample_neighbor(prev, row_ptr, col_idx, nnz);
                double u = (double)rand() / (double)RAND_MAX;
                i64 t = row[s - 2];
                if (x == t) {
                    if (u < prob_0) { selected = x; break; }
                } else if (csr_is_neighbor(x, t, row_ptr, col_idx)) {
                    if (u < prob_1) { selected = x; break; }
                } else if (u < prob_2) {
                    selected = x; break;
                }
            }
            row[s] = selected;
            prev = selected;
        }
    }
    return 0;
}

/* ------------------------------------------------------------ edge-list walk */

/* csrc/cpu/rw_cpu_edge_list.cpp:8-35.  nei = node_edge_index[N,2], inclusive [first,last]. */
static i64 el_sample_neighbor(i64 node, i64 jump, const i64 *nei, const i64 *el, i64 pad)
{
    if (node != pad) {
        i64 first = nei[2 * node], last = nei[2 * node + 1];
        if (first == -1 || last == -1) return pad;
        return el[2 * sample_int(first, last) + 1];
    }
    return jump;
}

/* csrc/cpu/rw_cpu_edge_list.cpp:37-62 -- half-open scan: the LAST out-edge is skipped. */
static int el_is_neighbor(i64 new_node, i64 previous_node, const i64 *nei, const i64 *el)
{
    i64 first = nei[2 * previous_node], last = nei[2 * previous_node + 1];
    if (first == -1 || last == -1) return 0;
    for (i64 i = first; i < last; i++)
        if (el[2 * i + 1] == new_node) return 1;
    return 0;
}

/* walk_edge_list_cpu, csrc/cpu/rw_cpu_edge_list.cpp:240-266; uniform :64-126; biased :128-238.
 * n_index_rows bounds the is_neighbor lookup of a padding previous node: the source reads row
 * `padding_idx` of node_edge_index, which is out of bounds when padding_idx == N; the oracle
 * treats an out-of-range row as "no out-edges" instead of reading past the tensor. */
int orc_walk_edge_list(const i64 *el, const i64 *nei, i64 n_index_rows, const i64 *targets,
                       i64 n_walks, double p, double q, int walk_length, int seed, i64 pad,
                       int restart, i64 *out)
{
    i64 ws = (i64)walk_length + 1;
    srand((unsigned)seed);
    if (p == 1.0 && q == 1.0) {
        for (i64 w = 0; w < n_walks; w++) {
            i64 *row = out + w * ws;
            i64 start = targets[w];
            i64 jump = restart ? start : pad;
            i64 prev = start;
            row[0] = start;
            for (i64 s = 1; s < ws; s++) {
                prev = el_sample_neighbor(prev, jump, nei, el, pad);
                row[s] = prev;
            }
        }
        return 0;
    }
    double prob_0, prob_1, prob_2;
    rejection_probs(p, q, &prob_0, &prob_1, &prob_2);
    for (i64 w = 0; w < n_walks; w++) {
        i64 *row = out + w * ws;
        i64 start = targets[w];
        i64 jump = restart ? start : pad;
        row[0] = start;
        if (ws < 2) continue;
        row[1] = el_sample_neighbor(start, jump, nei, el, pad);
        i64 prev = row[1];
        for (i64 s = 2; s < ws; s++) {
            i64 selected;
            for (;;) {
                i64 x = el_sample_neighbor(prev, jump, nei, el, pad);
                double u = (double)rand() / (double)RAND_MAX;
                i64 t = row[s - 2];
                /* rw_cpu_edge_list.cpp:203-230: the x==t test is a separate `if`, so a rejected
                 * return falls through into the chain below. */
                if (x == t) {
                    if (u < prob_0) { selected = x; break; }
                }
                if (x == pad) {
                    if (u < prob_0) { selected = jump; break; }
                } else if (t >= 0 && t < n_index_rows && el_is_neighbor(x, t, nei, el)) {
                    if (u < prob_1) { selected = x; break; }
                } else if (u < prob_2) {
                    selected = x; break;
                }
            }
            row[s] = selected;
            prev = selected;
        }
    }
    return 0;
}

/* --------------------------------------------------------------- triple walk */

/* walk_triples_cpu, csrc/cpu/rw_cpu_triples.cpp:105-127; uniform_walk_triples :48-103;
 * sample_neighbor :11-46.  rti = relation_tail_index[N,2]; out is [n_walks, 2*walk_length+1].
 * `restart` is accepted and ignored, as in the source. */
int orc_walk_triples(const i64 *triples, const i64 *rti, const i64 *targets, i64 n_walks,
                     int walk_length, i64 pad, int restart, int seed, i64 *out)
{
    (void)restart;
    i64 ws = 2 * (i64)walk_length + 1;
    srand((unsigned)seed);
    for (i64 w = 0; w < n_walks; w++) {
        i64 *row = out + w * ws;
        i64 prev = targets[w];
        row[0] = prev;
        for (i64 s = 1; s < ws; s += 2) {
            i64 rel = pad, tail = pad;
            if (prev != pad) {
                i64 first = rti[2 * prev], last = rti[2 * prev + 1];
                if (!(first == -1 || last == -1)) {
                    i64 k = sample_int(first, last);
                    rel = triples[3 * k + 1];
                    tail = triples[3 * k + 2];
                }
            }
            row[s] = rel;
            row[s + 1] = tail;
            prev = tail;
        }
    }
    return 0;
}

/* -------------------------------------------------------- skip-gram windows */

/* to_windows_cpu, csrc/cpu/windows_cpu.cpp:5-77.  walks [n_walks, wl]; outputs
 * target[K], pos[K, W-1], neg[K, W-1], K = n_walks*(wl-W+1). */
int orc_windows(const i64 *walks, i64 n_walks, i64 wl, int window_size, i64 num_nodes, int seed,
                i64 *target, i64 *pos, i64 *neg)
{
    srand((unsigned)seed);
    i64 W = window_size, mid = W / 2, step_end = wl - W + 1;
    for (i64 w = 0; w < n_walks; w++) {
        const i64 *walk = walks + w * wl;
        for (i64 s = 0; s < step_end; s++) {
            i64 k = w * step_end + s;
            target[k] = walk[s + mid];
            i64 j = 0;
            for (i64 i = 0; i < W; i++)
                if (i != mid) pos[k * (W - 1) + j++] = walk[s + i];
            for (i64 i = 0; i < W - 1; i++)
                neg[k * (W - 1) + i] = (i64)rand() % num_nodes;
        }
    }
    return 0;
}

/* to_windows_cbow_cpu, csrc/cpu/windows_cpu.cpp:80-159.  Outputs pos_nodes[K], neg_nodes[K],
 * windows[K, W-1]. */
int orc_windows_cbow(const i64 *walks, i64 n_walks, i64 wl, int window_size, i64 num_nodes,
                     int seed, i64 *pos_nodes, i64 *neg_nodes, i64 *windows)
{
    srand((unsigned)seed);
    i64 W = window_size, mid = W / 2, step_end = wl - W + 1;
    for (i64 w = 0; w < n_walks; w++) {
        const i64 *walk = walks + w * wl;
        for (i64 s = 0; s < step_end; s++) {
            i64 k = w * step_end + s;
            i64 pos_node = walk[s + mid];
            pos_nodes[k] = pos_node;
            i64 neg_node = sample_int(0, num_nodes - 1);
            int max_checks = 0;
            while (neg_node == pos_node && max_checks <= 100) {
                neg_node = sample_int(0, num_nodes - 1);
                max_checks++;
            }
            neg_nodes[k] = neg_node;
            i64 j = 0;
            for (i64 i = 0; i < W; i++)
                if (i != mid) windows[k * (W - 1) + j++] = walk[s + i];
        }
    }
    return 0;
}

/* ----------------------------------------------------------- triple windows */

/* Positive rows shared by to_windows_triples_cpu (csrc/cpu/windows_cpu.cpp:215-280) and
 * to_windows_triples_cbow_cpu (:389-456).  pw points at pos_windows[k] = [2W,3].
 * Left loop runs hop = 0..W INCLUSIVE (:216/:391); row W is then overwritten by the right loop.
 * The head slot of a left row holds walk[rel_idx] (the relation), :223-224, as pinned by the
 * reference golden [10,10,27] (tests/test_windows.py:150). */
static void triple_pos_rows(const i64 *walk, i64 wl, i64 r, i64 W, i64 pad, i64 *pw)
{
    for (i64 hop = 0; hop <= W && hop < 2 * W; hop++) {
        i64 rel_idx = r - (hop + 1) * 2, head_idx = rel_idx - 1, tail_idx = rel_idx + 1;
        pw[hop * 3 + 0] = head_idx >= 0 ? walk[rel_idx] : pad;
        pw[hop * 3 + 1] = rel_idx >= 0 ? walk[rel_idx] : pad;
        pw[hop * 3 + 2] = tail_idx >= 0 ? walk[tail_idx] : pad;
    }
    for (i64 hop = 0; hop < W; hop++) {
        i64 rel_idx = r + (hop + 1) * 2, head_idx = rel_idx - 1, tail_idx = rel_idx + 1;
        i64 *o = pw + (hop + W) * 3;
        o[0] = head_idx < wl ? walk[head_idx] : pad;
        o[1] = rel_idx < wl ? walk[rel_idx] : pad;
        o[2] = tail_idx < wl ? walk[tail_idx] : pad;
    }
}

/* to_windows_triples_cpu, csrc/cpu/windows_cpu.cpp:161-310.  Outputs target[K,3],
 * pos[K,2W,3], neg[K,2W,3], K = n_walks*((wl-1)/2).  num_nodes is unused by the source. */
int orc_windows_triples(const i64 *walks, i64 n_walks, i64 wl, int window_size, i64 num_nodes,
                        i64 pad, const i64 *triples, i64 n_triples, int seed, i64 *target,
                        i64 *pos, i64 *neg)
{
    (void)num_nodes;
    srand((unsigned)seed);
    i64 W = window_size, per_walk = (wl - 1) / 2;
    for (i64 w = 0; w < n_walks; w++) {
        const i64 *walk = walks + w * wl;
        i64 ti = 0;
        for (i64 r = 1; r < wl - 1; r += 2, ti++) {
            i64 k = per_walk * w + ti;
            target[k * 3 + 0] = walk[r - 1];
            target[k * 3 + 1] = walk[r];
            target[k * 3 + 2] = walk[r + 1];
            triple_pos_rows(walk, wl, r, W, pad, pos + k * 2 * W * 3);
            for (i64 hop = 0; hop < 2 * W; hop++) {
                i64 idx = sample_int(0, n_triples - 1);
                neg[(k * 2 * W + hop) * 3 + 0] = triples[idx * 3 + 0];
                neg[(k * 2 * W + hop) * 3 + 1] = triples[idx * 3 + 1];
                neg[(k * 2 * W + hop) * 3 + 2] = triples[idx * 3 + 2];
            }
        }
    }
    return 0;
}

/* to_windows_triples_cbow_cpu, csrc/cpu/windows_cpu.cpp:312-475.  Outputs pos_triples[K,3],
 * neg_triples[K,3], pos_windows[K,2W,3].  The negative is drawn BEFORE the window rows (:363). */
int orc_windows_triples_cbow(const i64 *walks, i64 n_walks, i64 wl, int window_size,
                             i64 num_nodes, i64 pad, const i64 *triples, i64 n_triples, int seed,
                             i64 *pos_triples, i64 *neg_triples, i64 *pos_windows)
{
    (void)num_nodes;
    srand((unsigned)seed);
    i64 W = window_size, per_walk = (wl - 1) / 2;
    for (i64 w = 0; w < n_walks; w++) {
        const i64 *walk = walks + w * wl;
        i64 ti = 0;
        for (i64 r = 1; r < wl - 1; r += 2, ti++) {
            i64 k = per_walk * w + ti;
            i64 ph = walk[r - 1], pr = walk[r], pt = walk[r + 1];
            pos_triples[k * 3 + 0] = ph;
            pos_triples[k * 3 + 1] = pr;
            pos_triples[k * 3 + 2] = pt;
            i64 idx = sample_int(0, n_triples - 1);
            i64 nh = triples[idx * 3], nr = triples[idx * 3 + 1], nt = triples[idx * 3 + 2];
            int max_checks = 0;
            while (nh == ph && nr == pr && nt == pt && max_checks <= 100) {
                idx = sample_int(0, n_triples - 1);
                nh = triples[idx * 3]; nr = triples[idx * 3 + 1]; nt = triples[idx * 3 + 2];
                max_checks++;
            }
            neg_triples[k * 3 + 0] = nh;
            neg_triples[k * 3 + 1] = nr;
            neg_triples[k * 3 + 2] = nt;
            triple_pos_rows(walk, wl, r, W, pad, pos_windows + k * 2 * W * 3);
        }
    }
    return 0;
}
