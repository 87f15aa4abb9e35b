"""Independent numpy restatement of the compact fractal-error form (include/vrdd.h, vrdd_error_entry):
per chunk of 32 consecutive voxels, round k = the k-th error of every voxel with NE > k, in voxel order."""
import numpy as np

CHUNK = 32
DTYPE = [("bin", "<i4"), ("value", "<f4")]


def round_major(codebook, errors_dense):
    cb = np.asarray(codebook).reshape(-1, 4)
    ed = np.asarray(errors_dense, dtype=np.float32).reshape(cb.shape[0], -1, 2)
    out = []
    for c0 in range(0, cb.shape[0], CHUNK):
        ne = cb[c0:c0 + CHUNK, 3]
        for k in range(int(ne.max(initial=0))):
            for v in np.nonzero(ne > k)[0]:
                out.append((int(ed[c0 + v, k, 0]), ed[c0 + v, k, 1]))
    return np.array(out, dtype=DTYPE)


def chunk_offsets(codebook):
    ne = np.asarray(codebook).reshape(-1, 4)[:, 3].astype(np.uint64)
    cum = np.concatenate([[0], np.cumsum(ne)]).astype(np.uint64)
    n = ne.shape[0]
    return np.concatenate([cum[0:n:CHUNK], cum[-1:]])
