"""Checkpoint hand-off with the reference: ``train_acc.py:492-503`` writes ``{"model", "ema", "opt", "args"}`` with
``state_dict()`` keys (``module.``-prefixed under DDP / accelerate) and ``sample_acc.py:70-78`` loads
``state_dict["ema"]`` after stripping that prefix.  ``zigma_b200.ZigMa`` keeps the reference's parameter names and
shapes, so a reference checkpoint (e.g. the HF ``taohu/zigma`` ``.pt`` files, README.md:154-159) loads as is and a
checkpoint written here loads in the reference."""
import torch


def strip_module_prefix(state_dict):
    return {k.replace("module.", ""): v for k, v in state_dict.items()}      # sample_acc.py:73 (replace, not removeprefix)


def load_reference_checkpoint(model, path_or_dict, which="ema", strict=True):
    """Loads ``which`` ("ema" as sample_acc.py does, or "model") of a reference-format checkpoint into ``model``;
    a bare state dict is accepted too.  Returns the (missing, unexpected) key lists of ``load_state_dict``."""
    ck = path_or_dict
    if not isinstance(ck, dict):
        ck = torch.load(ck, map_location="cpu", weights_only=False)
    sd = ck[which] if which in ck and isinstance(ck[which], dict) else ck
    out = model.load_state_dict(strip_module_prefix(sd), strict=strict)
    eng = getattr(model, "_engine", None)
    if eng is not None:
        eng.refresh()          # the sampling engine caches packed weights and CUDA graphs
    return out


def save_reference_checkpoint(path, model, ema_model=None, opt_state=None, args=None, ddp_prefix=False, train_steps=0, best_fid=None):
    """Writes the dictionary train_acc.py:492-503 writes ({"model", "ema", "opt", "args", "train_steps", "best_fid"}).
    ``opt_state``: ``optimizer.state_dict()`` -- ``FusedAdamWEMA.state_dict()`` emits torch.optim.AdamW's layout."""
    pre = (lambda sd: {"module." + k: v for k, v in sd.items()}) if ddp_prefix else (lambda sd: dict(sd))
    ck = {"model": pre(model.state_dict()), "ema": pre((ema_model or model).state_dict()), "opt": opt_state, "args": args,
          "train_steps": int(train_steps), "best_fid": best_fid}
    torch.save(ck, path)
    return ck
