"""The flexible-block chain against the REFERENCE'S OWN device code.

tests/golden/ref_gpu_flex_v1.npz holds what the kernels of dataProcessing() (/root/reference/volumeRender_kernel.cu:
892-1796) and d_render's queryMethod 8 / 9 / 0 computed on a B200, compiled where they lie (oracle/Makefile `ref`,
ref_driver64) and driven by oracle/ref_driver.cu in mode "flexscrub" on the synthetic span store of
tools/ref_pin_flex.py (tests/flex_synth.py, seed 17, 64^3 raw volume, padded to the table sizes initCuda hard-codes).

What the reference's build defines here (tools/ref_pin_flex.py has the measurements):
  * flexibleFractalDecoding() returns a pointer to its local array (:224-251); nvcc keeps the stores of original[0..54]
    (unflipped) resp. original[63..8] (flipped) only; the other bins are read uninitialised — zero on scrubbed local
    memory, which ref_pin_flex.as_the_reference_build_decodes models with a doubled template table;
  * from its second wave on a CTA of d_querySpanNew inherits the `decoded` arrays other spans left in local memory, so
    only the first wave is deterministic: every block of the 2x2x2 configuration (block size 32) with its frames, and
    blocks 0..17 of the reference's own configuration (block size 6).
On that the source's algorithm — span decomposition, hash/linear span look-up, flip/shift, error merge with clamp,
weights, normalisation, the sign pattern of the eight corners, the un-normalised statistics, un-normalised sampling of the
zero-padded block volume — reproduces the binary to float rounding, and the frames of queryMethod 8 byte for byte."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.fixture(scope="module")
def flexpin():
    import ref_pin_flex as P
    fx = dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_gpu_flex_v1.npz")))
    assert int(fx["seed"]) == P.SEED and int(fx["raw"]) == P.RAW and tuple(fx["block_sizes"]) == P.BLOCKS
    out = {"fx": fx, "P": P}
    for b in P.BLOCKS:
        t = P.tables(b)
        out[b] = (t, P.as_the_reference_build_decodes(t))
    return out


def _byte_diff(a, b):
    return np.abs(np.ascontiguousarray(a).view(np.uint8).astype(np.int16) - np.ascontiguousarray(b).view(np.uint8).astype(np.int16))


@pytest.mark.parametrize("block", [32, 6])
def test_oracle_chain_matches_the_reference_binary(oracle, flexpin, block):
    fx = flexpin["fx"]
    intent_tables, modelled_tables = flexpin[block]
    ref = fx[f"blocks_b{block}"]
    pinned = int(fx[f"pinned_b{block}"])
    mine, nb, missing = oracle.flex_process(modelled_tables, block)
    assert missing == 0 and list(fx[f"dims_b{block}"]) == [nb[0] * nb[1] * nb[2], nb[0], nb[1], nb[2]]
    assert pinned == (8 if block == 32 else 18)
    np.testing.assert_allclose(mine[:pinned, :3], ref[:pinned, :3], rtol=1e-6, atol=0)
    assert not ref[:pinned, 3].any()
    # the source's intent (all 64 bins kept; what the oracle and the kernels implement) is a different number: the
    # synthetic templates are dense, so the nine / eight dropped bins carry mass in every span
    intent, _, _ = oracle.flex_process(intent_tables, block)
    assert (np.abs(intent[:pinned, 0] - ref[:pinned, 0]) / ref[:pinned, 0] > 1e-4).all()


def test_oracle_frames_of_query_methods_8_9_0_match_the_reference_binary(oracle, flexpin):
    fx = flexpin["fx"]
    mine, nb, _ = oracle.flex_process(flexpin[32][1], 32)
    views, image = fx["views_b32"], tuple(int(v) for v in fx["image"])
    for k in range(views.shape[0]):
        for j, qm in enumerate((8, 9, 0)):
            img, _ = oracle.render_flex(mine, nb, views[k], image=image, query_method=qm)
            d = _byte_diff(img, fx["images_b32"][k, j])
            assert d.max() <= 1 and (d != 0).sum() <= 16, (k, qm, int(d.max()), int((d != 0).sum()))
        assert (fx["images_b32"][k, 0] != 0).sum() > 20000                    # queryMethod 8 (entropy) is a real picture;
        assert not fx["images_b32"][k, 1:].any()                             # 9 / 0: un-normalised mean / variance saturate the TF


@pytest.mark.gpu
@pytest.mark.parametrize("block", [32, 6])
def test_cuda_chain_matches_the_reference_binary(renderer, flexpin, block):
    """csrc/flex.cu through the C ABI on the span store as the reference's build decodes it."""
    import torch
    import vrdd_b200 as V
    fx = flexpin["fx"]
    ref = fx[f"blocks_b{block}"]
    pinned = int(fx[f"pinned_b{block}"])
    r = renderer
    r.flex_set_tables_host(flexpin[block][1])
    assert r.flex_process(block) == 0
    got, gdims = r.flex_get_blocks_host()
    assert list(fx[f"dims_b{block}"][1:]) == list(gdims)
    # corner sums are reduced with warp shuffles instead of the reference's unordered atomics: rounding only
    np.testing.assert_allclose(got[:pinned, :3], ref[:pinned, :3], rtol=2e-5, atol=2e-4)
    if block == 32:
        views, (w, h) = fx["views_b32"], (int(v) for v in fx["image"])
        for k in range(views.shape[0]):
            r.set_view(views[k])
            for j, qm in enumerate((8, 9, 0)):
                out = torch.zeros(h, w, dtype=torch.int32, device="cuda")
                r.render(out, w, h, V.default_render_params(query_method=qm), clear_misses=False)
                r.synchronize()
                d = _byte_diff(out.cpu().numpy().view(np.uint32), fx["images_b32"][k, j])
                assert d.max() <= 1 and (d != 0).sum() <= 64, (k, qm, int(d.max()), int((d != 0).sum()))
