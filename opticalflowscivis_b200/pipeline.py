"""Host -> device -> host streaming around `Model.inference` (the data edge of the hot path).

The reference copies each pair to the GPU, runs the model and copies the result back on one stream
(`Flow-3D/train.py:257-268`, `Flow-2D/train.py:279-312`).  On a B200 the 256^3 interpolation itself takes a few
milliseconds, so the PCIe copies (33 MB of uint8 in, 67 MB of fp32 out per pair) would cost a third of the end-to-end time
if they were serialised with it.  `StreamedInterpolator` keeps three CUDA streams busy instead — upload of pair i+1,
compute of pair i, download of pair i-1 — with double-buffered device inputs and pinned host outputs; the copies then
hide completely behind the kernels (both PCIe directions run concurrently with compute on their own copy engines).

No arithmetic happens here: the uint8 -> fp32 `/255` conversion is the reference loader's (`Datasets/read_data.py`,
`Flow-3D/load_datasets.py`), done on the device after the upload so that only bytes cross PCIe.
"""
from __future__ import annotations

import os
from typing import Callable, Iterable, Iterator, Optional, Sequence, Tuple

import torch


def read_raw_volume(path: str, shape: Sequence[int] = (256, 256, 256), out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One time step of the droplet ensemble: a headerless uint8 file, x fastest (`README.md:24-25`), read exactly like
    `np.fromfile(path, dtype='uint8').resize(256, 256, 256)` (`Datasets/read_data.py:116-119`) but straight into PINNED host
    memory, shaped (1, 1, *shape) for `StreamedInterpolator`.  A short file is zero-padded like ndarray.resize does, a longer
    one is truncated."""
    n = 1
    for d in shape:
        n *= int(d)
    if out is None:
        out = torch.zeros((1, 1) + tuple(int(d) for d in shape), dtype=torch.uint8)
        if torch.cuda.is_available():
            out = out.pin_memory()
    elif out.dtype != torch.uint8 or out.numel() != n or not out.is_contiguous():
        raise ValueError("read_raw_volume: `out` must be a contiguous uint8 tensor of the requested size")
    view = memoryview(out.view(-1).numpy())
    with open(path, "rb") as f:
        got = f.readinto(view)
    if got < n:
        out.view(-1)[got:] = 0
    return out


def raw_pairs(paths: Sequence[str], shape: Sequence[int] = (256, 256, 256), stride: int = 2) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
    """(volume[i], volume[i + stride]) pairs of a sorted file list — the inputs whose middle time step the model
    reconstructs (`Flow-3D/load_datasets.py`: every other member is held out) — as pinned uint8 tensors."""
    paths = sorted(paths, key=os.path.basename)
    for i in range(0, len(paths) - stride, stride):
        yield read_raw_volume(paths[i], shape), read_raw_volume(paths[i + stride], shape)


class StreamedInterpolator:
    """Runs `model.inference(img0, img1)` over a sequence of HOST pairs with copy/compute overlap.

    model      : rife.Model2D / Model3D (or anything with `.inference(img0, img1)`)
    to_float   : device-side conversion of an uploaded batch to the fp32 tensor the model expects
                 (default: uint8 -> x/255, fp32 passes through)
    select     : picks the tensor to download from the model's return value (default: the interpolated frame/volume)
    """

    def __init__(self, model, device: Optional[torch.device] = None, to_float: Optional[Callable] = None,
                 select: Optional[Callable] = None, depth: int = 2, out_u8: bool = False):
        """out_u8: download the result as bytes, `(merged * 255).byte()` computed on the device — the reference's own export
        (Flow-3D/inference_img.py:105) — instead of fp32: a quarter of the D2H bytes.  Off by default (fp32 like
        `Model.inference`)."""
        self.model = model
        self.out_u8 = out_u8
        self.dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        from . import ops
        self.to_float = to_float or (lambda t: ops.u8_to_f32(t, 255.0) if t.dtype == torch.uint8 else t)
        self.select = select or (lambda out: out[0] if torch.is_tensor(out[0]) else out[0][2])
        self.depth = max(2, depth)
        self.s_in = torch.cuda.Stream(self.dev)
        self.s_out = torch.cuda.Stream(self.dev)
        self._slots = []
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _slot(self, i, h0):
        while len(self._slots) <= i:
            self._slots.append({"x0": torch.empty(h0.shape, dtype=h0.dtype, device=self.dev),
                                "x1": torch.empty(h0.shape, dtype=h0.dtype, device=self.dev),
                                "up": torch.cuda.Event(), "free": torch.cuda.Event(), "done": torch.cuda.Event(),
                                "res": None, "out": None})
        return self._slots[i]

    def run(self, pairs: Iterable[Tuple[torch.Tensor, torch.Tensor]], outs: Optional[Iterable[torch.Tensor]] = None) -> Iterator[torch.Tensor]:
        """pairs: iterable of (img0, img1) PINNED host tensors of one fixed shape/dtype; outs: optional iterable of pinned
        host tensors receiving the results (allocated on demand otherwise).  Yields the host results in order; a yielded
        tensor is complete (its download has been synchronised) when it is handed out."""
        from . import ops
        comp = torch.cuda.current_stream(self.dev)
        outs_it = iter(outs) if outs is not None else None
        pending = []
        k = 0
        for h0, h1 in pairs:
            sl = self._slot(k % self.depth, h0)
            with torch.cuda.stream(self.s_in):
                if k >= self.depth:
                    self.s_in.wait_event(sl["free"])          # compute of the pair that used this slot has consumed it
                sl["x0"].copy_(h0, non_blocking=True)
                sl["x1"].copy_(h1, non_blocking=True)
                sl["up"].record(self.s_in)
            self.h2d_bytes += 2 * h0.numel() * h0.element_size()
            comp.wait_event(sl["up"])
            x0, x1 = self.to_float(sl["x0"]), self.to_float(sl["x1"])
            res = self.select(self.model.inference(x0, x1))
            if self.out_u8:
                res = ops.f32_to_u8(res, 255.0)
            elif getattr(self.model, "_graphs", None) is not None:
                # graph mode: `res` is the graph's own output buffer, which the NEXT replay overwrites while this pair's download
                # may still be reading it (record_stream cannot protect a graph-private buffer) — hand the download a copy
                res = res.clone()
            sl["free"].record(comp)                            # (a pass-through to_float hands the slot itself to the model)
            ev = torch.cuda.Event()
            ev.record(comp)
            out = next(outs_it) if outs_it is not None else torch.empty(res.shape, dtype=res.dtype).pin_memory()
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev)
                out.copy_(res, non_blocking=True)
                res.record_stream(self.s_out)
                done = torch.cuda.Event()
                done.record(self.s_out)
            self.d2h_bytes += res.numel() * res.element_size()
            pending.append((out, done))
            k += 1
            while len(pending) > self.depth:
                o, d = pending.pop(0)
                d.synchronize()
                yield o
        for o, d in pending:
            d.synchronize()
            yield o
