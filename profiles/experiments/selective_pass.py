"""CPU emulation of the split-bf16 tensor-core arithmetic, per GEMM of the message-passing edge chain.

For each of the four edge-sized GEMMs (edge_affine.0, edge_affine.2, theta_edge.1, theta_edge.3) the operand
split is emulated in torch CPU fp32:  x = hi + lo, hi = bf16(x), lo = bf16(x - hi);
    passes 3 : A_hi B_hi + A_lo B_hi + A_hi B_lo     (the shipped "bf16x3" mode)
    passes 2a: A_hi B_hi + A_lo B_hi                 (activations split, weights rounded)
    passes 2b: A_hi B_hi + A_hi B_lo                 (weights split, activations rounded)
    passes 1 : A_hi B_hi
Every other operation is the fp32 oracle.  Output: force error of LJ-258 / TIP3P-774 (random-init weights)
against the all-fp32 oracle, as max|err|/max|F| and max|err|/rms(F - mean F).

    python profiles/experiments/selective_pass.py > profiles/experiments/selective_pass_r02.txt
"""
import os, sys
import numpy as np, torch
import torch.nn.functional as Fnn
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import model as om, neighbor as onb
from gamd_b200.weights import random_state_dict, water_bonds

FIX = os.path.join(ROOT, "tests", "golden", "fixtures")
NAMES = ["edge_affine.mlp_layer.0", "edge_affine.mlp_layer.2", "theta_edge.mlp_layer.1", "theta_edge.mlp_layer.3"]


def split(x):
    hi = x.to(torch.bfloat16).float()
    lo = (x - hi).to(torch.bfloat16).float()
    return hi, lo


def lin_emul(x, w, b, mode):
    if mode == "fp32":
        return Fnn.linear(x, w, b)
    xh, xl = split(x)
    wh, wl = split(w)
    y = xh @ wh.T
    if mode in ("3", "2a"):
        y = y + xl @ wh.T
    if mode in ("3", "2b"):
        y = y + xh @ wl.T
    return y + b


def mp_layer(sd, l, h, e, center, neigh, modes):
    p = f"graph_conv.conv.{l}."
    L = lambda n, x, m="fp32": lin_emul(x, sd[p + n + ".weight"], sd[p + n + ".bias"], m)
    hn = om._ln(sd, f"graph_conv.norm_layers.{l}", h)
    edge_code = L(NAMES[1], Fnn.silu(L(NAMES[0], e, modes[0])), modes[1])
    a = edge_code + L("src_affine", hn)[neigh] + L("dst_affine", hn)[center]
    m = L(NAMES[3], Fnn.silu(L(NAMES[2], Fnn.silu(a), modes[2])), modes[3])
    agg = torch.zeros_like(hn)
    agg.index_add_(0, center, hn[neigh] * m)
    return L("phi.mlp_layer.1", Fnn.silu(L("phi_dst", hn) + L("phi_edge", agg))) + h


@torch.no_grad()
def forward(sd, kind, pos, edge, box, x, bond, modes):
    c, n = edge[0].long(), edge[1].long()
    flag = om.bond_flags(bond, c, n, pos.shape[0]) if kind != "lj" else None
    e = om.edge_encode(sd, om.edge_features(sd, pos, c, n, box, flag, True))
    h = sd["node_emb"].repeat((pos.shape[0], 1)) if kind == "lj" else om._lin(sd, "node_encoder", x)
    for l in range(om.n_conv_layers(sd)):
        h = mp_layer(sd, l, h, e, c, n, modes)
    return om.decode(sd, h).numpy().astype(np.float64)


def system(kind):
    if kind == "lj":
        pos = np.load(os.path.join(FIX, "lj_init_pos.npy")); box, rc = 27.27, 7.5
        sd = random_state_dict(1, 5.2, 1.5, kind="lj"); x = None; bond = None
    else:
        pos = np.load(os.path.join(FIX, "water_init_pos.npy")); box, rc = 20.0, 4.2
        sd = random_state_dict(4, 2.9, 0.9, kind="water"); bond = water_bonds(258)
        x = torch.zeros(774, 1); x[::3] = 1.0
    sd = {k: torch.as_tensor(v) for k, v in sd.items()}
    pw = onb.wrap_f32(pos.astype(np.float32), box)
    edge = torch.from_numpy(onb.edges_bruteforce(pw, box, rc))
    return sd, torch.from_numpy(pw), edge, box, x, bond


def main():
    torch.set_num_threads(8)
    rows = [("all fp32", ["fp32"] * 4), ("all 3-pass (shipped bf16x3)", ["3"] * 4), ("all 1-pass (shipped bf16)", ["1"] * 4),
            ("all 2a (A split)", ["2a"] * 4), ("all 2b (B split)", ["2b"] * 4)]
    for g in range(4):
        for m in ("2a", "2b", "1"):
            modes = ["3"] * 4
            modes[g] = m
            rows.append((f"3-pass except {NAMES[g]} -> {m}", modes))
    for kind in ("lj", "water"):
        sd, pos, edge, box, x, bond = system(kind)
        ref = forward(sd, kind, pos, edge, box, x, bond, ["fp32"] * 4)
        rms = np.sqrt(((ref - ref.mean(0)) ** 2).mean())
        print(f"== {kind}: N={pos.shape[0]} E={edge.shape[1]} max|F|={np.abs(ref).max():.4f} rms(F-mean)={rms:.4f}")
        for name, modes in rows:
            out = forward(sd, kind, pos, edge, box, x, bond, modes)
            d = np.abs(out - ref).max()
            print(f"{name:58s} rel-to-max {d / np.abs(ref).max():.2e}   rel-to-rms {d / rms:.2e}")


if __name__ == "__main__":
    main()
