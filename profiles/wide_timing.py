"""Time one force evaluation of the dynamic-box water model through the reference-facing call
(``WaterMDDynamicBoxNet.forward(pos_lst, x, box_size_lst, cutoff)``) on the water fixture (774 atoms, 20 A box,
cutoff 5.0 A ~ the 9.5 bohr of code/water/test_script/test_nosehoover_hb.py:75): the 256 / 128 / 256 x 5 DFT-water
shape on the generic-width fp32 kernels (csrc/model_wide.cu) beside the 128-wide shape on the tensor-core kernels."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gamd_b200 import _capi
from gamd_b200.nn_module import WaterMDDynamicBoxNet
from gamd_b200.weights import random_state_dict

fix = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "fixtures")
pos = torch.as_tensor(np.mod(np.load(os.path.join(fix, "water_init_pos.npy")), 20.0).astype(np.float32), device="cuda:0")
x = torch.zeros(774, 1, device="cuda:0"); x[::3] = 1.0
box = np.array([20.0, 20.0, 20.0], dtype=np.float32)
out = {}
for name, (D, H, De, L), prec in (("w256_fp32_generic", (256, 128, 256, 5), _capi.PREC_FP32),
                                  ("w128_bf16x3_tcgen05", (128, 128, 128, 4), _capi.PREC_BF16X3),
                                  ("w128_fp32_cudacore", (128, 128, 128, 4), _capi.PREC_FP32)):
    m = WaterMDDynamicBoxNet(1, D, 3, hidden_dim=H, conv_layer=L, edge_embedding_dim=De, drop_edge=False, use_layer_norm=True)
    m.load_state_dict(random_state_dict(6, 2.9, 0.9, kind="dynbox", use_bond=False, encoding_size=D, hidden_dim=H,
                                        edge_embedding_dim=De, conv_layer=L))
    m.cuda().eval()
    m.context(precision=prec)
    for _ in range(3):
        m([pos], x, [box], 5.0)
    ne = m._ctx.neighbor_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    n = 30
    for _ in range(n):
        m([pos], x, [box], 5.0)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / n
    flop_edge = 2 * ((64 * H + H * H + H * De) + L * (De * 128 + 128 * H + H * H + H * D))
    out[name] = dict(ms_per_force_eval=ms, n_edges=ne, atoms=774, gflops_edge_mlps=flop_edge * ne / ms / 1e6)
    print(name, json.dumps(out[name]))
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "wide_timing.json"), "w"), indent=1)
