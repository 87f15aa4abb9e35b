"""rsqrt.approx.f32 of the B200 on [4, 6] as a fixture for the oracle (tests/golden/rsqrt_approx_b200_v1.npz).

    python tools/make_rsqrt_table.py dump <raw.u32>      runs tools/rsqrt_dump on the GPU box
    python tools/make_rsqrt_table.py pack <raw.u32>      raw dump -> the fixture: per float of [4, 6], the distance in
                                                         ulps (int8) from (float)(1.0 / sqrt((double)x)), the value
                                                         the oracle computes without the table

The eye ray of d_render is normalised with rsqrtf (helper_math.h normalize(), volumeRender_kernel.cu:295); the
argument u*u + v*v + 4 lies in [4, 6].  With the table the oracle's "reference build" rounding
(Oracle.set_reference_build) reproduces the ray set-up of the reference's own binary bit for bit."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LO, HI = 0x40800000, 0x40C00000
FIXTURE = os.path.join(ROOT, "tests", "golden", "rsqrt_approx_b200_v1.npz")


def dump(path):
    exe = os.path.join(ROOT, "tools", "rsqrt_dump")
    subprocess.run([exe, path, f"{LO:08X}", f"{HI:08X}"], check=True)


def pack(path):
    got = np.fromfile(path, np.uint32)
    assert got.size == HI - LO + 1, got.size
    x = np.arange(LO, HI + 1, dtype=np.uint32).view(np.float32)
    ref = (1.0 / np.sqrt(x.astype(np.float64))).astype(np.float32).view(np.uint32)
    delta = got.astype(np.int64) - ref.astype(np.int64)
    assert np.abs(delta).max() <= 4, int(np.abs(delta).max())
    print("ulps from the correctly rounded value:", {int(k): int((delta == k).sum()) for k in np.unique(delta)})
    np.savez_compressed(FIXTURE, lo_bits=np.array(LO, np.uint32), hi_bits=np.array(HI, np.uint32), delta=delta.astype(np.int8),
                        gpu=np.array("NVIDIA B200, rsqrt.approx.f32"))
    print("fixture:", FIXTURE, os.path.getsize(FIXTURE), "bytes")


if __name__ == "__main__":
    {"dump": dump, "pack": pack}[sys.argv[1]](sys.argv[2])
