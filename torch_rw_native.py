"""`import torch_rw_native` compatibility: the reference's extension module name (setup.py:55-56),
served by torch_random_walk_b200.native (ctypes over libtrw_b200.so)."""
from torch_random_walk_b200.native import (walk, walk_edge_list, walk_triples, to_windows, to_windows_cbow,  # noqa: F401
                                           to_windows_triples, to_windows_triples_cbow)
