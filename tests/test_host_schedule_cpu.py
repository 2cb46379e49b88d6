"""The schedule of the host-buffer entry points (rtat_b200_{d,s}gemm_host, rtat_b200_mg_dgemm_host) is pure host
arithmetic (rtatblas_b200/csrc/host_schedule.h) exported through rtat_b200_host_schedule, so its invariants are
checked here without a GPU: column blocks tile n exactly, K-panels tile k in order, the leading blocks are a prefix,
and the schedules DESIGN.md quotes for the BASELINE configs are the ones the library produces."""
import itertools

import pytest

import rtatblas_b200 as rt


def _check(sc, n, k):
    blocks, panels, lead = sc["blocks"], sc["panels"], sc["lead"]
    # column blocks: contiguous, in order, cover [0, n); all but the tail are multiples of 128 columns
    pos = 0
    for c0, cnt in blocks:
        assert c0 == pos and cnt > 0
        pos += cnt
    assert pos == n
    assert all(c0 % 128 == 0 for c0, _ in blocks)
    assert 1 <= len(blocks) <= 14  # 8 blocks, the last one halved (multi-GPU caller) or tapered down to 256 columns
    # K-panels: contiguous, in order, cover [0, k); starts are multiples of 16 (the DGEMM's k-tile)
    pos = 0
    for k0, kc in panels:
        assert k0 == pos and kc > 0 and k0 % 16 == 0
        pos += kc
    assert pos == k
    assert len(panels) <= 24  # 8 panels, the first one halved — or a head that grows from 256 k to the regular width
    # the K-panel phase exists only with more than one panel, and then for at least one and at most all blocks
    if len(panels) == 1:
        assert lead == 0
    else:
        assert 1 <= lead <= len(blocks)


@pytest.mark.parametrize("elem", [8, 4])
def test_schedule_invariants_over_a_sweep(elem):
    for m, n, k, world in itertools.product([1, 300, 4097, 8192, 32768], [1, 129, 256, 703, 3001, 8192, 16384],
                                            [1, 77, 2047, 2049, 4096, 8192, 32768], [0, 1, 2, 8]):
        for transa in (0, 1):
            lda = (k if transa else m) + 3
            _check(rt.host_schedule(elem, transa, m, n, k, lda, world), n, k)


def test_small_or_shallow_problems_take_the_plain_column_block_pipeline():
    # A under 32 MiB, or k < 2048: one panel (all of A first), no leading blocks
    assert rt.host_schedule(8, 0, 1000, 900, 700, 1000)["lead"] == 0
    sc = rt.host_schedule(8, 1, 1000, 703, 3000, 3000)   # 24 MB of A
    assert sc["lead"] == 0 and sc["panels"] == [(0, 3000)] and [c for _, c in sc["blocks"]] == [256, 256, 191]
    assert len(rt.host_schedule(8, 0, 8192, 8192, 2047, 8192)["panels"]) == 1
    # a single column block cannot be pipelined at all
    assert rt.host_schedule(8, 0, 8192, 200, 8192, 8192)["blocks"] == [(0, 200)]


def test_baseline_config_schedules_match_the_design_notes():
    # config 2 (8192^3, op TN): four leading blocks would balance product and upload exactly (growth 0.07), so it takes a
    # fifth (growth 0.24) and a head growing from 256 k; 10 column blocks (the last of 8 tapered down to 256 columns)
    sc = rt.host_schedule(8, 1, 8192, 8192, 8192, 8192)
    widths = [kc for _, kc in sc["panels"]]
    assert sc["lead"] == 5 and widths[0] == 256 and sum(widths) == 8192 and len(widths) <= 13
    assert all(w <= 256 + 0.25 * sum(widths[:i]) + 16 for i, w in enumerate(widths[:-1]))
    assert all(b >= a for a, b in zip(widths[:-2], widths[1:-1])) and widths[-1] >= 256
    assert [c for _, c in sc["blocks"]] == [1024] * 7 + [512, 256, 256]   # tapered tail: only 256 columns' copy back is exposed
    # the override keeps round 2's first schedule reachable: 8 panels of 1024 with the first one halved, 4 leading blocks
    sc = rt.host_schedule(8, 1, 8192, 8192, 8192, 8192, panels=8, lead=4)
    assert sc["panels"] == [(0, 512), (512, 512)] + [(1024 * i, 1024) for i in range(1, 8)] and sc["lead"] == 4
    # config 3 (4097 x 3001 x 2049, op NT): the transfer of A dominates the upload, but the copy back of C (98 MB) outlasts the
    # whole product, so only two blocks lead (a leading block is complete, and its copy back can start, only when all of A is up)
    sc = rt.host_schedule(8, 0, 4097, 3001, 2049, 4097)
    assert sc["lead"] == 2 and len(sc["blocks"]) == 8 and len(sc["panels"]) == 3
    # config 5 on one GPU: one leading block whose product is 1.4x its panel's upload -> a head growing from 256 k by 0.37 of
    # what is already up (no longer capped at the regular 4096), and a tail tapered from 4096 columns down to 256
    sc = rt.host_schedule(8, 0, 32768, 32768, 32768, 32768)
    widths = [kc for _, kc in sc["panels"]]
    assert sc["lead"] == 1 and widths[0] == 256 and max(widths) <= 4 * 4096 and sum(widths) == 32768
    assert all(b >= a for a, b in zip(widths[:-2], widths[1:-1])) and len(widths) <= 20
    assert all(w <= 256 + 0.38 * sum(widths[:i]) + 16 for i, w in enumerate(widths[:-1]))
    assert [c for _, c in sc["blocks"]][-5:] == [2048, 1024, 512, 256, 256]
    # config 5 per rank (m = k = 32768, n = 32768 / world): the multi-GPU caller's margin on the leading blocks
    assert rt.host_schedule(8, 0, 32768, 16384, 32768, 32768, world=2)["lead"] == 1
    assert rt.host_schedule(8, 0, 32768, 8192, 32768, 32768, world=4)["lead"] == 2
    assert rt.host_schedule(8, 0, 32768, 4096, 32768, 32768, world=8)["lead"] == 3


def test_overrides_and_bad_arguments():
    sc = rt.host_schedule(8, 1, 8192, 8192, 8192, 8192, panels=4, lead=100)
    assert sc["lead"] == len(sc["blocks"]) and len(sc["panels"]) == 5
    with pytest.raises(rt.RtatError):
        rt.host_schedule(2, 0, 10, 10, 10, 10)
    with pytest.raises(rt.RtatError):
        rt.host_schedule(8, 0, -1, 10, 10, 10)


def test_multi_gpu_panels_do_not_depend_on_the_ranks_column_count():
    """rtat_b200_mg_dgemm_host is collective and n_loc may differ by one column between ranks (n = 256 * world + 1 puts one
    rank on either side of the one-block / two-block threshold): every rank must cut K into the same panels, because the
    peers' arrival flags are indexed by panel."""
    for world in (2, 4, 8):
        for transa in (0, 1):
            for m, k in [(8192, 8192), (4096, 2048), (32768, 32768), (8192, 2047), (1000, 3000)]:
                lda = k if transa else m
                seen = set()
                for n_loc in (1, 200, 256, 257, 511, 512, 513, 1024, 4096):
                    sc = rt.host_schedule(8, transa, m, n_loc, k, lda, world)
                    seen.add(tuple(sc["panels"]))
                    _check(sc, n_loc, k)
                assert len(seen) == 1, f"world {world} transa {transa} m {m} k {k}: panel schedules differ across n_loc: {seen}"
