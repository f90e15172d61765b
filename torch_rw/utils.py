from torch_random_walk_b200.utils import (to_csr, nodes_tensor, to_edge_list_indexed, build_node_edge_index,  # noqa: F401
                                          build_relation_tail_index, csr_from_edge_index)
