"""TEST INFRASTRUCTURE ONLY.  The deterministic synthetic weights / inputs (pure numpy + torch, no kernels) live in
zigma_b200/synth.py so that the product can build its benchmark model without importing the oracle.  The file is loaded here
BY PATH, not through the package: importing ``oracle`` (e.g. ``bench.py --impl reference``) must not import ``zigma_b200``."""
import importlib.util
import os

_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "zigma_b200", "synth.py")
_spec = importlib.util.spec_from_file_location("_zigma_synth_by_path", _path)
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)

_rs = _mod._rs
synth_param = _mod.synth_param
synth_state_dict = _mod.synth_state_dict
synth_latents = _mod.synth_latents
synth_scan_inputs = _mod.synth_scan_inputs
__all__ = ["synth_param", "synth_state_dict", "synth_latents", "synth_scan_inputs"]
