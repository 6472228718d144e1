"""Re-export: the deterministic synthetic weights / inputs live in zigma_b200/synth.py (pure
numpy/torch, no kernels) so that bench.py can build its model without importing the oracle."""
from zigma_b200.synth import *  # noqa: F401,F403
from zigma_b200.synth import _rs  # noqa: F401
