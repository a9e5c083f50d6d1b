"""The fixed service order of the message-passing pair kernel (csrc/mp_tc2cta.cu, STATIC) and of the edge encoder
(csrc/enc_tc.cu, STATIC), restated on the host: the MMA warp issues GEMM n into TMEM block (3 + n) & 3 and every epilogue
thread derives its accumulator block from (stage, slot, slots active in its group) - MP kernel - or from its own GEMM
count - encoder.  This checks, for every tail shape, that the two sides agree and that a GEMM never writes a block that
holds a live operand (its own or another slot's).  CPU only: pure index arithmetic."""
import pytest

NSLOT = 3


def leader_sequence(n_groups, nact_last, stages):
    """(slot, stage, group, D block, A block) per GEMM, as the leader's loop issues them."""
    home = list(range(NSLOT))
    n = 0
    out = []
    for i in range(n_groups):
        nact = nact_last if i == n_groups - 1 else NSLOT
        for s in range(stages):
            for g in range(NSLOT):
                if g >= nact:
                    continue
                blk = (3 + n) & 3
                out.append((g, s, i, blk, home[g], nact))
                home[g] = blk
                n += 1
    return out


@pytest.mark.parametrize("n_groups", [1, 2, 3, 7])
@pytest.mark.parametrize("nact_last", [1, 2, 3])
def test_mp_kernel_blocks(n_groups, nact_last):
    live = {g: g for g in range(NSLOT)}         # slot -> block holding its current operand (initial homes 0, 1, 2)
    for g, s, i, blk, a_blk, nact in leader_sequence(n_groups, nact_last, 4):
        # epilogue side (mp_tc2cta.cu, GAMD_STAGE): block of stage s of slot g in a group with `nact` slots at work
        assert blk == (3 + s * nact + g) & 3
        assert a_blk == live[g]                                  # the GEMM reads the slot's operand where it was left
        # the written block is neither this slot's operand nor the operand of any other slot still at work; slots that
        # are absent in the tail group hold nothing live: the leader waits for their "last sums have been read" arrival
        # before the tail group's first GEMM (mp_tc2cta.cu: "absent slot of the tail group")
        active_others = [live[q] for q in range(NSLOT) if q != g and q < nact]
        assert blk != a_blk and blk not in active_others
        live[g] = blk
        if s == 3 and nact == NSLOT:
            assert blk == g                                       # full groups: a tile ends where it began


@pytest.mark.parametrize("n_groups", [1, 2, 5])
@pytest.mark.parametrize("nact_last", [1, 2, 3])
def test_encoder_blocks(n_groups, nact_last):
    nseq = {g: g for g in range(NSLOT)}         # enc_tc.cu: per-thread counter, += nact per stage
    live = {g: g for g in range(NSLOT)}
    for g, s, i, blk, a_blk, nact in leader_sequence(n_groups, nact_last, 3):
        mine = (3 + nseq[g]) & 3
        nseq[g] += nact
        assert mine == blk and a_blk == live[g]
        active_others = [live[q] for q in range(NSLOT) if q != g and q < nact]
        assert blk != a_blk and blk not in active_others
        live[g] = blk
