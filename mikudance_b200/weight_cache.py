"""Packed fp16 weight cache.

The reference assembles the denoising UNet's weights at every start-up from three checkpoints
(`from_pretrained_2d`: SD-1.5 `diffusion_pytorch_model.*` + the motion-module file, src/models/unet_3d_mix.py:639-683;
then `denoising_unet.pth`, scripts/inference_video.py:111-114) and converts them to fp16 on the way to the GPU.
UNetEngine then re-lays them out for the kernels (OIHW -> [Cout, (kh kw Cin)], q|k|v concatenation, GEGLU panel
interleave, the pe @ Wq^T tables, the stacked time-embedding projection).  This module stores that final, packed
form in ONE safetensors file so a later start-up is a single read + host->device copy:

* transparent cache: with MDK_WEIGHT_CACHE=<dir>, `UNetEngine` looks up `<dir>/mdk-packed-<key>.safetensors` before it
  packs, and writes the file after it packed (the reference UNet's RefUNetEngine likewise).  <key> is a position-sensitive content hash of the model's live
  state dict (computed on the device the weights are on) + the engine class + the layout version + the GEGLU panel width, so a changed
  checkpoint, a changed kernel layout or a different build never hits a stale file.
* explicit packed checkpoint: `UNet3DConditionModel.save_packed(path)` / `.from_packed(path, device)` — the file
  also carries the constructor kwargs; `from_packed` builds the module tree on the `meta` device (no parameter
  storage) and the engine straight from the file, skipping the three-file merge entirely.

File format: safetensors; tensors named by their path in the engine ("down.0.res.1.w1"); metadata = {"layout":
version, "key", "geglu_block", "skeleton": JSON of the engine's packed attribute tree, "ctor": JSON kwargs}.
"""
from __future__ import annotations

import hashlib
import json
import os
from typing import Dict, Optional

import torch

LAYOUT_VERSION = 1          # bump whenever UNetEngine._pack changes what it produces


class _Node:
    """Attribute bag of the engine's packed tree (same role as engine._Resnet)."""


# ------------------------------------------------------------------------------------------------
# content hash of a model's weights
# ------------------------------------------------------------------------------------------------
def _tensor_digest(t: torch.Tensor) -> torch.Tensor:
    """Two int64 words per tensor, position-sensitive, computed where the tensor lives.  Integer sums wrap
    mod 2^64, so the result does not depend on the reduction order (bit-identical on CPU and GPU)."""
    x = t.detach().contiguous().reshape(-1)
    if x.element_size() % 2 == 0:
        x = x.view(torch.int16)
    else:
        x = x.view(torch.uint8)
    x = x.to(torch.int64)
    n = x.numel()
    row = 1024
    pad = (-n) % row
    if pad:
        x = torch.cat([x, x.new_zeros(pad)])
    x = x.view(-1, row)
    wcol = torch.arange(1, row + 1, dtype=torch.int64, device=x.device)
    r = (x * wcol).sum(1)
    wrow = (torch.arange(r.numel(), dtype=torch.int64, device=x.device) % 8191) * 2 + 1
    return torch.stack([(r * wrow).sum(), x.sum() + n])


def weights_key(model, tag: str = "") -> str:
    """Hex key of the model's state dict (names, shapes, dtypes and contents) + `tag`."""
    h = hashlib.blake2b(digest_size=16)
    h.update(tag.encode())
    digs = []
    for name, t in sorted(model.state_dict().items()):
        h.update(f"{name}|{tuple(t.shape)}|{t.dtype}\n".encode())
        digs.append(_tensor_digest(t))
    if digs:
        h.update(torch.stack(digs).cpu().numpy().tobytes())
    return h.hexdigest()


def layout_tag(engine) -> str:
    return f"{type(engine).__name__}|layout{LAYOUT_VERSION}|geglu{engine.geglu_block}|"


# ------------------------------------------------------------------------------------------------
# engine tree <-> (skeleton, tensors)
# ------------------------------------------------------------------------------------------------
def _encode(v, path: str, tensors: Dict[str, torch.Tensor], engine, modnames):
    if isinstance(v, torch.Tensor):
        tensors[path] = v
        return {"t": path}
    if isinstance(v, torch.nn.Module):
        return {"m": modnames[id(v)]}
    if v is None or isinstance(v, (bool, int, float, str)):
        return {"v": v}
    if isinstance(v, (list, tuple)):
        return {"l" if isinstance(v, list) else "u": [_encode(x, f"{path}.{i}", tensors, engine, modnames)
                                                      for i, x in enumerate(v)]}
    if isinstance(v, dict):
        for side in ("down", "up"):                      # block-plan entries: stored by reference
            for i, d in enumerate(engine.plan[side]):
                if v is d:
                    return {"p": [side, i]}
        raise TypeError(f"weight cache: unexpected dict at {path}")
    if hasattr(v, "__dict__"):
        return {"o": {k: _encode(x, f"{path}.{k}", tensors, engine, modnames) for k, x in vars(v).items()}}
    raise TypeError(f"weight cache: cannot store {type(v).__name__} at {path}")


def _decode(e, tensors, engine):
    (kind, val), = e.items()
    if kind == "t":
        return tensors[val]
    if kind == "m":
        return engine.model.get_submodule(val)
    if kind == "v":
        return val
    if kind == "l":
        return [_decode(x, tensors, engine) for x in val]
    if kind == "u":
        return tuple(_decode(x, tensors, engine) for x in val)
    if kind == "p":
        return engine.plan[val[0]][val[1]]
    if kind == "o":
        n = _Node()
        for k, x in val.items():
            setattr(n, k, _decode(x, tensors, engine))
        return n
    raise ValueError(f"weight cache: bad skeleton node {kind!r}")


def export_packed(engine):
    """(tensors, skeleton) of everything `UNetEngine._pack` produced."""
    modnames = {id(m): n for n, m in engine.model.named_modules()}
    tensors: Dict[str, torch.Tensor] = {}
    skel = {a: _encode(getattr(engine, a), a, tensors, engine, modnames) for a in engine._packed_attrs}
    return tensors, skel


def import_packed(engine, tensors, skel) -> None:
    for a, e in skel.items():
        setattr(engine, a, _decode(e, tensors, engine))
    engine._packed_attrs = sorted(skel)


# ------------------------------------------------------------------------------------------------
# files
# ------------------------------------------------------------------------------------------------
def write_file(engine, path: str, key: str, ctor: Optional[dict] = None) -> None:
    from safetensors.torch import save_file
    tensors, skel = export_packed(engine)
    meta = {"layout": str(LAYOUT_VERSION), "key": key, "geglu_block": str(engine.geglu_block),
            "skeleton": json.dumps(skel), "ctor": json.dumps(ctor) if ctor is not None else ""}
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    tmp = f"{path}.tmp.{os.getpid()}"
    # host copies: a packed tensor may alias a module parameter (safetensors refuses shared storage)
    def host(v):
        v = v.detach().contiguous()
        return v.cpu() if v.device.type != "cpu" else v.clone()
    save_file({k: host(v) for k, v in tensors.items()}, tmp, metadata=meta)
    os.replace(tmp, path)          # atomic: concurrent ranks write identical content


def read_file(engine, path: str, want_key: Optional[str] = None) -> bool:
    """Fill `engine` from `path`.  False (engine untouched) when the file does not match this build's layout,
    the GEGLU panel width or `want_key`."""
    from safetensors import safe_open
    with safe_open(path, framework="pt", device=str(engine.dev)) as f:
        meta = f.metadata() or {}
        if meta.get("layout") != str(LAYOUT_VERSION) or meta.get("geglu_block") != str(engine.geglu_block):
            return False
        if want_key is not None and meta.get("key") != want_key:
            return False
        tensors = {k: f.get_tensor(k) for k in f.keys()}
    import_packed(engine, tensors, json.loads(meta["skeleton"]))
    return True


def read_ctor(path: str) -> dict:
    from safetensors import safe_open
    with safe_open(path, framework="pt", device="cpu") as f:
        meta = f.metadata() or {}
    if not meta.get("ctor"):
        raise RuntimeError(f"{path}: not a packed checkpoint written by UNet3DConditionModel.save_packed "
                           "(a MDK_WEIGHT_CACHE entry carries no constructor arguments)")
    return json.loads(meta["ctor"])


class PackedWeightCache:
    def __init__(self, directory: str):
        self.dir = directory
        self.last = None            # "hit" | "miss" | "stale" (tests, logging)

    @staticmethod
    def from_env() -> Optional["PackedWeightCache"]:
        d = os.environ.get("MDK_WEIGHT_CACHE", "")
        return PackedWeightCache(d) if d else None

    def path_for(self, key: str) -> str:
        return os.path.join(self.dir, f"mdk-packed-{key}.safetensors")

    def load(self, engine) -> Optional[str]:
        """Returns the key; engine is filled on a hit (self.last == "hit")."""
        key = weights_key(engine.model, layout_tag(engine))
        path = self.path_for(key)
        self.last = "miss"
        if os.path.isfile(path):
            self.last = "hit" if read_file(engine, path, key) else "stale"
        return key

    def store(self, engine, key: str) -> str:
        path = self.path_for(key)
        write_file(engine, path, key)
        return path
