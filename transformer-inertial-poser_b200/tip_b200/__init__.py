"""tip_b200: B200-native (sm_100a) implementation of the Transformer-Inertial-Poser hot path.

``TF_RNN_Past_State`` mirrors the reference class; ``capi`` is the ctypes binding of
``include/tip_b200.h``; ``build`` compiles ``libtip_b200.so`` in-tree.
"""
from .module import TF_RNN_Past_State, state_dict_keys  # noqa: F401
