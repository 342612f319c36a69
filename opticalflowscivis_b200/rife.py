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
        self.device()

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

    def update(self, *a, **k):
        raise NotImplementedError("Model.update (training step) is the next tier of the hot path — SURVEY.md §8(f).1")

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

    def inference(self, img0, img1, scale_list=[4, 2, 1], TTA=False, timestep=0.5):
        """Flow-2D/model/RIFE.py:66-78 -> (merged[3], flow_list[3], mask_list[3]); TTA returns the flip-averaged frame."""
        flow, mask, merged, *_ = self._run(img0, img1, scale_list, timestep, only_last=False)
        if not TTA:
            return merged, flow, mask
        _, _, merged2, *_ = self._run(img0.flip(2).flip(3), img1.flip(2).flip(3), scale_list, timestep, only_last=True)
        return (merged[2] + merged2[2].flip(2).flip(3)) / 2


class Model3D(_ModelBase):
    ND = 3

    def inference(self, img0, img1, scale_list=[4, 2, 1], TTA=False, timestep=0.5):
        """Flow-3D/model/RIFE.py:67-79 -> (merged[2], flow_list[3], mask_list[2]); TTA is 'not implemented' upstream."""
        if TTA:
            raise NotImplementedError("the reference 3-D Model.inference prints 'not implemented' for TTA (RIFE.py:77)")
        flow, mask, merged, *_ = self._run(img0, img1, scale_list, timestep, only_last=True)
        return merged[2], flow, mask
