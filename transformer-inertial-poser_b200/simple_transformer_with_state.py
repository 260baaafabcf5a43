"""Drop-in for the reference module of the same name
(/root/reference/simple_transformer_with_state.py): put this directory ahead of the reference
checkout on ``sys.path`` and ``from simple_transformer_with_state import TF_RNN_Past_State``
(offline_testing_simple.py:80, live_demo_new.py:16) resolves to the B200-native class."""
from tip_b200.module import TF_RNN_Past_State  # noqa: F401
