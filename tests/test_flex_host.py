"""The two stages of the reference's flexible-block chain that the reference itself pins with known answers
(SURVEY.md §4): the block partition of d_divideBlock and the bitwise prefix-span decomposition of
d_queryBlockNew.  Host functions of libvrdd.so; the rest of that chain is not built."""
import numpy as np


def _divide(V, dims, block):
    n = V.lib().vrdd_flex_divide_blocks(dims[0], dims[1], dims[2], block, None, 0)
    out = np.zeros((n, 6), np.int32)
    assert V.lib().vrdd_flex_divide_blocks(dims[0], dims[1], dims[2], block, out.ctypes.data, n) == n
    return out


def test_changelog_known_answer_block_13():
    """ver1.9.6.txt:77 — with block size 30 on the 64^3 volume, block 13 is spanLow(31,31,31) spanHigh(60,60,60)."""
    import vrdd_b200 as V
    spans = _divide(V, (64, 64, 64), 30)
    assert spans.shape[0] == 27
    assert spans[13].tolist() == [31, 31, 31, 60, 60, 60]
    assert spans[0].tolist() == [1, 1, 1, 30, 30, 30] and spans[26].tolist() == [61, 61, 61, 64, 64, 64]
    assert spans[1].tolist() == [31, 1, 1, 60, 30, 30]                      # x fastest (:1016-1026)


def test_slide_14_known_answer_prefix_25():
    """presentation.pdf p.14 — 25 = 11001b -> [25,25], [17,24], [1,16] (volumeRender_kernel.cu:1248-1259)."""
    import vrdd_b200 as V
    out = np.zeros((32, 2), np.int32)
    n = V.lib().vrdd_flex_prefix_spans(25, out.ctypes.data)
    assert n == 3 and out[:3].tolist() == [[25, 25], [17, 24], [1, 16]]


def test_partition_and_decomposition_properties():
    import vrdd_b200 as V
    for dims, block in (((64, 64, 64), 6), ((64, 64, 64), 64), ((50, 40, 30), 7)):   # 6 is dataProcessing()'s own (:1737)
        spans = _divide(V, dims, block)
        vol = np.zeros(dims[::-1], np.int32)
        for lx, ly, lz, hx, hy, hz in spans:
            vol[lz - 1:hz, ly - 1:hy, lx - 1:hx] += 1
        assert np.all(vol == 1)                                                   # a partition: every voxel once
    out = np.zeros((32, 2), np.int32)
    for x in list(range(0, 70)) + [255, 256, 1023, 2 ** 20 + 5]:
        n = V.lib().vrdd_flex_prefix_spans(x, out.ctypes.data)
        assert n == bin(x).count("1")
        pieces = out[:n]
        assert sum(int(h - l + 1) for l, h in pieces) == x                        # they tile [1, x]
        for l, h in pieces:
            size = int(h - l + 1)
            assert size & (size - 1) == 0 and (l - 1) % size == 0                 # power of two, aligned
        if n:
            assert pieces[0][1] == x and pieces[-1][0] == 1
