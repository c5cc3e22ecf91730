# Namespace shim so that `from nbodyhpc import kdtree` / `from nbodyhpc.kdtree import KDTree`
# (the reference's import path, kdtree/setup.py:113-125) resolves to the B200-native package.
