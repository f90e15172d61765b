"""Import-compatibility shim: `from torch_rw import rw, utils` resolves to torch_random_walk_b200,
so code written against the reference package (README.md:16-43 there) runs unchanged."""
