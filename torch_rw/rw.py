from torch_random_walk_b200.rw import *  # noqa: F401,F403
from torch_random_walk_b200.rw import (walk, walk_edge_list, walk_triples, to_windows, to_windows_cbow,  # noqa: F401
                                       to_windows_triples, to_windows_triples_cbow)
