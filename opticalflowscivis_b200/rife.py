"""`Model` wrapper shared by the 2-D and 3-D packages — mirrors Flow-2D/model/RIFE.py:19-78 and
Flow-3D/model/RIFE.py:18-79 (constructor, train/eval/device, load_model/save_model, inference)."""
from __future__ import annotations

import torch

from .ifnet import IFNet


class _ModelBase:
    ND = 0

    def __init__(self, local_rank=-1, arbitrary=False, precision="bf16", engine="auto"):
        if arbitrary:
            # Flow-*/model/IFNet_m.py is upstream RGB code never adapted to 1-channel data and never enabled by any
            # driver: out of scope.
            raise NotImplementedError("arbitrary-timestep IFNet_m is dead code in the reference and is not provided")
        if not torch.cuda.is_available():
            raise RuntimeError("opticalflowscivis_b200 needs a CUDA device (the reference hard-codes cuda too: RIFE.py:16-17)")
        self.flownet = IFNet(self.ND, precision=precision, engine=engine)
        self.local_rank = local_rank
        self._graphs = None
        self.device()

    # ---------------------------------------------------------------------------------------------- CUDA graphs (opt-in)
    def enable_cuda_graphs(self, on=True):
        """Replay `inference` from a CUDA graph captured per (input shape, scale_list): every launch of the call (~45, each
        with a ctypes call and, for the convolutions, a tensor-map encode) costs host time, and small problems are bound by
        it — one 160x224 pair takes 1.28 ms eagerly (all of it host enqueue time) and 0.45 ms replayed; 256^3 volumes gain
        1-2 % (tests/graph_probe.py).  Off by default because it changes aliasing, not numerics: the inputs are copied
        into buffers owned by the graph and the returned tensors are the graph's output buffers, OVERWRITTEN by the next
        `inference` call with the same shapes (clone what you keep).  Graphs are dropped when a parameter changes."""
        self._graphs = {} if on else None
        return self

    def enable_training_graph(self, on=True):
        """Replay forward + backward of `update` from a CUDA graph captured per input shape (train.Trainer.forward_backward_graphed).
        Same numerics; the returned `merged` / info tensors are then the graph's buffers, overwritten by the next `update`."""
        self._train_graph = bool(on)
        return self

    def _param_signature(self):
        return tuple(p._version for p in self.flownet.parameters()) + tuple(p.data_ptr() for p in self.flownet.parameters())

    def _graphed(self, fn, img0, img1, scale_list):
        for t, name in ((img0, "img0"), (img1, "img1")):
            if not t.is_cuda:
                raise TypeError(f"{name}: expected a CUDA tensor (no CPU path)")
        sig = self._param_signature()
        if self._graphs.get("sig") != sig:
            self._graphs.clear()
            self._graphs["sig"] = sig
        key = (tuple(img0.shape), tuple(img1.shape), img0.dtype, tuple(scale_list), img0.device.index)
        ent = self._graphs.get(key)
        if ent is None:
            a, b = img0.detach().clone().contiguous(), img1.detach().clone().contiguous()
            side = torch.cuda.Stream(device=img0.device)
            side.wait_stream(torch.cuda.current_stream(img0.device))
            with torch.cuda.stream(side), torch.no_grad():      # warm-up off the capture: workspaces, packed weights, attributes
                for _ in range(2):
                    fn(a, b)
            torch.cuda.current_stream(img0.device).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g), torch.no_grad():
                out = fn(a, b)
            ent = self._graphs[key] = (g, a, b, out)
        g, a, b, out = ent
        a.copy_(img0)
        b.copy_(img1)
        g.replay()
        return out

    def train(self):
        self.flownet.train()

    def eval(self):
        self.flownet.eval()

    def device(self):
        dev = torch.device("cuda", self.local_rank if self.local_rank >= 0 else torch.cuda.current_device())
        self.flownet.to(dev)

    def load_model(self, model_name, path, rank=0):
        """The reference keeps only DDP-prefixed keys ('module.', RIFE.py:45-57) and so only round-trips under DDP;
        here both prefixed and plain `state_dict`s are accepted."""
        if rank <= 0:
            sd = torch.load("{}/{}".format(path, model_name), map_location="cpu")
            sd = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in sd.items()}
            self.flownet.load_state_dict(sd)
            print("loaded {}".format(model_name))

    def save_model(self, model_name, path, rank=0):
        if rank == 0:
            torch.save(self.flownet.state_dict(), "{}/{}".format(path, model_name))
            print("saved {}".format(model_name))


    def _run(self, img0, img1, scale_list, timestep, only_last):
        for t, name in ((img0, "img0"), (img1, "img1")):
            if not t.is_cuda:
                raise TypeError(f"{name}: expected a CUDA tensor (no CPU path)")
        self.flownet.only_last = only_last
        try:
            return self.flownet.forward_pair(img0, img1, scale_list, timestep=timestep)
        finally:
            self.flownet.only_last = False


class Model2D(_ModelBase):
    ND = 2

    def update(self, imgs, gt, dataset="droplet2d", learning_rate=0, mul=1, training=True, flow_gt=None):
        """Flow-2D/model/RIFE.py:80-336 on the 1-channel dataset branch (droplet2d / vimeo2d): LapLoss student + teacher,
        0.01 * distillation, 1e-5 * photometric, backward, AdamW (opticalflowscivis_b200/train.py)."""
        from . import train
        return train.update(self, imgs, gt, learning_rate, mul, training, flow_gt, dataset=dataset)

    def inference(self, img0, img1, scale_list=[4, 2, 1], TTA=False, timestep=0.5):
        """Flow-2D/model/RIFE.py:66-78 -> (merged[3], flow_list[3], mask_list[3]); TTA returns the flip-averaged frame."""
        if self._graphs is not None and not TTA:
            return self._graphed(lambda a, b: self._eager2d(a, b, scale_list, timestep), img0, img1, scale_list)
        return self._eager2d(img0, img1, scale_list, timestep, TTA)

    def _eager2d(self, img0, img1, scale_list, timestep, TTA=False):
        flow, mask, merged, *_ = self._run(img0, img1, scale_list, timestep, only_last=False)
        if not TTA:
            return merged, flow, mask
        _, _, merged2, *_ = self._run(img0.flip(2).flip(3), img1.flip(2).flip(3), scale_list, timestep, only_last=True)
        return (merged[2] + merged2[2].flip(2).flip(3)) / 2


class Model3D(_ModelBase):
    ND = 3

    def update(self, imgs, gt, learning_rate=0, mul=1, training=True, flow_gt=None):
        """Flow-3D/model/RIFE.py:81-275: forward with the teacher block, L1 + L1(teacher) + 0.1 * distillation, backward, AdamW
        (opticalflowscivis_b200/train.py).  Returns (merged[2], info dict with the reference's keys)."""
        from . import train
        return train.update(self, imgs, gt, learning_rate, mul, training, flow_gt)

    def inference(self, img0, img1, scale_list=[4, 2, 1], TTA=False, timestep=0.5):
        """Flow-3D/model/RIFE.py:67-79 -> (merged[2], flow_list[3], mask_list[2]); TTA is 'not implemented' upstream."""
        if TTA:
            raise NotImplementedError("the reference 3-D Model.inference prints 'not implemented' for TTA (RIFE.py:77)")
        if self._graphs is not None:
            return self._graphed(lambda a, b: self._eager3d(a, b, scale_list, timestep), img0, img1, scale_list)
        return self._eager3d(img0, img1, scale_list, timestep)

    def _eager3d(self, img0, img1, scale_list, timestep):
        flow, mask, merged, *_ = self._run(img0, img1, scale_list, timestep, only_last=True)
        return merged[2], flow, mask
