"""Launch configuration the stacked tcgen05 conv engine picks for every conv layer of a 3-D IFNet inference (no GPU work):
usage: describe_layers.py [size=256] [pairs=4]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opticalflowscivis_b200 import _C, ifnet

s = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
net = ifnet.IFNet(3)
L = _C.lib()
buf = ctypes.create_string_buffer(512)
for bi, (blk, sc) in enumerate(zip((net.block0, net.block1, net.block2), (4, 2, 1))):
    layers = list(blk.layers())
    layers[11] = blk._heads_shuffle
    layers[0] = blk._s2d0
    if blk._s2d1_ok:
        layers[1] = blk._s2d1
    sp = (s // sc,) * 3
    for li, lay in enumerate(layers):
        d, osp = lay.desc(n, sp, _C.BF16, hfast=(li == 11))
        _C.check(L.ofsv_conv_halo_describe(ctypes.byref(d), buf, 512))
        print(f"block{bi} layer{li:2d} in {sp} Cin_s={lay.cin_s:3d} Cout_w={lay.cout_w:3d}: {buf.value.decode()}")
        sp = osp
