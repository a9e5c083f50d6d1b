"""Stage-level parity (SURVEY.md section 4 "kernel parity"): every intermediate of the MDNet forward against the
oracle's, so that a compensating error between the edge encoder and the message-passing layers cannot hide behind
a passing end-to-end force test.

The C ABI exposes the layer boundary through the domain-decomposition entry points (gamd_dd_begin = neighbor search
+ edge encoder + layer-0 node prologue; gamd_dd_layer(l) = one message-passing layer) and the scratch buffers
through gamd_debug_ptr.  Checked per arithmetic mode (fp32 CUDA-core anchor, tcgen05 bf16x3):

    e        edge-encoder output after LayerNorm, [E,128]      (nn_module.py:646)
    hn_l     LN_l(h_l), the layer's node input                  (nn_module.py:202)
    agg_l    sum over the receiver's edges of hn[src] * e_emb   (nn_module.py:142)
    h_{l+1}  the layer's output incl. the residual              (nn_module.py:147, :202)

Tolerances are absolute on O(1) quantities (LayerNorm outputs) and relative to max|.| elsewhere; the measured values
go to gpurun_out/parity_metrics.jsonl."""
import os

import numpy as np
import pytest
import torch

from gamd_b200 import _capi
from oracle import model as omodel
from helpers import FIX, make_ctx, record

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
# (e, hn, agg, h) relative to max |oracle value| of the stage
STAGE_TOL = {_capi.PREC_FP32: (5e-6, 5e-6, 2e-5, 2e-5), _capi.PREC_BF16X3: (1e-4, 1e-4, 2e-4, 2e-4)}


def decode_blob(blob_u8, n_edges):
    """bf16 hi/lo blob of the tensor-core path -> fp32 [E,128]: [tile][hi|lo][16 k-chunks][128 rows][8 bf16]."""
    nt = (n_edges + 127) // 128
    b = blob_u8[:nt * 65536].view(torch.int16).view(nt, 2, 16, 128, 8).to(torch.int32)
    f = (b << 16).view(torch.float32)                           # bf16 -> fp32 by bit placement
    e = (f[:, 0] + f[:, 1]).permute(0, 2, 1, 3).reshape(nt * 128, 128)   # [tile, row, kchunk, 8] -> [edge, 128]
    return e[:n_edges]


@pytest.mark.parametrize("prec", [_capi.PREC_FP32, _capi.PREC_BF16X3])
@pytest.mark.parametrize("system", ["lj258", "lj2064"])
def test_stage_level_parity(prec, system):
    pos = np.load(os.path.join(FIX, "lj_init_pos.npy")).astype(np.float64)
    box = 27.27
    if system == "lj2064":      # 8 images + jitter: many tiles, rows straddling tile and 32-edge block boundaries
        shifts = np.array([[i, j, k] for i in range(2) for j in range(2) for k in range(2)], dtype=np.float64) * 27.27
        rng = np.random.Generator(np.random.PCG64(9))
        pos = np.concatenate([pos + s for s in shifts]) + 0.05 * rng.standard_normal((2064, 3))
        box = 54.54
    n = len(pos)
    ctx, sd = make_ctx("lj", 1, 5.2, 1.5, scaler="scaler_lj.npz", precision=prec, max_atoms=n, max_edges=n * 40)
    L = len([k for k in sd if k.startswith("graph_conv.norm_layers.") and k.endswith(".weight")])
    ctx.dd_begin(torch.as_tensor(pos, device=DEV), n, box, 7.5)
    ctx.check_async_errors()
    ne = ctx.neighbor_count()
    perm = ctx.debug_tensor("perm", torch.int32, (n,)).long().cpu()
    col = ctx.debug_tensor("col_idx", torch.int32, (ne,)).long().cpu()
    dst = ctx.debug_tensor("edge_dst", torch.int32, (ne,)).long().cpu()
    row_ptr = ctx.debug_tensor("row_ptr", torch.int32, (n + 1,)).long().cpu()
    # the oracle on the SAME edge order (CSR order, caller ids) and the same wrapped fp32 positions the facade uses
    center, neigh = perm[dst], perm[col]
    p32 = torch.from_numpy(np.mod(pos, box)).float()
    _, inter = omodel.forward(sd, "lj", [p32], [torch.stack([center, neigh])], box, return_intermediates=True)
    te, thn, tagg, th = STAGE_TOL[prec]
    pname = {_capi.PREC_FP32: "fp32", _capi.PREC_BF16X3: "bf16x3"}[prec]

    def cmp(stage, got, want, tol):
        got, want = got.double().numpy(), want.double().numpy()
        err = np.abs(got - want).max() / np.abs(want).max()
        record("stage:" + system + ":" + stage, precision=pname, rel_to_max=err, tol=tol)
        print(f"{system} [{pname}] {stage}: max|err|/max|.| = {err:.3e} (tol {tol:g})")
        assert err <= tol, (stage, err, tol)

    if prec == _capi.PREC_FP32:
        e_gpu = ctx.debug_tensor("e_emb", torch.float32, (ne, 128)).cpu()
    else:
        nt = (ne + 127) // 128
        e_gpu = decode_blob(ctx.debug_tensor("e_emb", torch.uint8, (nt * 65536,)).cpu(), ne)
    cmp("e", e_gpu, inter["e"], te)
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(n)
    for l in range(L):
        hn = ctx.debug_tensor("hn", torch.float32, (n, 128)).cpu()[inv]      # LN_l(h_l), written by the node prologue
        cmp(f"hn{l}", hn, inter["hn"][l], thn)
        ctx.dd_layer(l)
        ctx.check_async_errors()
        agg = ctx.debug_tensor("agg", torch.float32, (n, 128)).cpu()
        want_agg = inter["agg"][l]
        if prec == _capi.PREC_FP32:
            # the fp32 anchor kernel stores only receiver rows whose edge run lies inside ONE 64-edge tile; rows that
            # straddle tiles are assembled from partial sums inside the node kernel and never reach the agg buffer
            whole = (row_ptr[1:] > row_ptr[:-1]) & ((row_ptr[:-1] // 64) == ((row_ptr[1:] - 1) // 64))
            assert whole.float().mean() > 0.5
            cmp(f"agg{l}", agg[whole], want_agg[perm[whole]], tagg)
        else:
            cmp(f"agg{l}", agg[inv], want_agg, tagg)
        if l + 1 < L:       # the last node update feeds the decoder from registers and does not store h
            h = ctx.debug_tensor("h", torch.float32, (n, 128)).cpu()[inv]
            cmp(f"h{l + 1}", h, inter["h"][l + 1], th)
    pred = ctx.debug_tensor("pred", torch.float32, (n, 3)).cpu()[inv]
    want = omodel.decode({k: torch.as_tensor(v) for k, v in sd.items()}, inter["h"][L])
    cmp("force_normalised", pred, want, 2e-5 if prec == _capi.PREC_FP32 else 1e-4)
    ctx.close()
