"""Host logic of the chunked transition store (sampler.TransitionBuffers) on CPU tensors: whatever the number of
chunks, the window `[H + K]` handed to the kernels must look like a store that is rolled before every launch."""
import numpy as np
import pytest
import torch

import msacl_b200  # noqa: F401
from msacl_b200.sampler import TransitionBuffers
from msacl_b200.specs import get_spec


@pytest.mark.parametrize("chunks", [1, 2, 3, 4])
@pytest.mark.parametrize("K,n_step", [(3, 5), (8, 5), (4, 1), (2, 9)])
def test_window_view_matches_rolled_store(chunks, K, n_step):
    spec, n = get_spec("DuctedFan"), 7
    tr = TransitionBuffers(spec, n, K, n_step, torch.device("cpu"), chunks=chunks)
    assert tr.M == chunks and tr.H == n_step - 1
    H = tr.H
    ref = {k: np.zeros((H + K,) + tuple(v.shape[1:]), dtype=v.numpy().dtype) for k, v in tr.fields().items()}
    rng = np.random.default_rng(0)
    for launch in range(3 * chunks + 2):
        tr.roll_history()
        for k in ref:                                        # reference: roll, then write the K new slices
            ref[k][:H] = ref[k][K:K + H].copy()
        cur = tr.fields()
        for k, v in cur.items():
            assert v.shape[0] == H + K
            new = rng.integers(0, 200, size=(K,) + tuple(v.shape[1:])).astype(ref[k].dtype)
            v[H:].copy_(torch.as_tensor(new))
            ref[k][H:] = new
        for k, v in tr.fields().items():                     # history + new slices of the current window
            assert np.array_equal(v.numpy(), ref[k]), (launch, k)
        d = tr.desc(H)                                       # the descriptor points at the first new slice
        assert d.obs == tr.obs[H:].data_ptr() and d.emit == tr.emit[H:].data_ptr()
    assert tr.obs.data_ptr() == tr.fields()["obs"].data_ptr()
