"""B200-native matrix-free operator apply for the libCEED + PETSc solid-mechanics mini-app.

Product code: `csrc/` (the `/gpu/b200` libCEED backend: C front-end + sm_100a CUDA kernels,
built into `libceed_b200.so`), `ceed.py` (ctypes binding of the libCEED user API it exports),
`mesh.py` / `matops.py` / `halo.py` (the PETSc-side harness: box meshes, MatShell callbacks,
halo exchange).  Nothing in this package imports `oracle/`.
"""
__version__ = "0.1.0"
