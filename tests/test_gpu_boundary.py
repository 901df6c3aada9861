"""Device-boundary format steps through the C ABI: the capture callback's stereo fold
`a + b` (devices.rs:244-262) and the playback callback's mono -> stereo duplicate (devices.rs:443-500).
Bit-exact against the same arithmetic in numpy; ragged lengths, device-pointer and host-pointer forms."""
import ctypes

import numpy as np
import pytest

from dsp_stuff_b200 import signals as S
from dsp_stuff_b200.engine import MEM_DEVICE, Engine
from tests.util import assert_bit_exact

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("C,n", [(2, 128), (3, 1), (5, 1001), (64, 4096 + 3), (1, 0)])
def test_fold_and_dup_bit_exact(C, n):
    e = Engine(C, block=128, max_samples=128)
    st = S.noise(C, 2 * n, seed=5).reshape(C, n, 2) if n else np.zeros((C, 0, 2), np.float32)
    mono = e.fold_stereo(st)
    assert_bit_exact(mono, st[:, :, 0] + st[:, :, 1], "stereo fold")
    back = e.dup_stereo(mono)
    assert back.shape == (C, n, 2)
    assert_bit_exact(back[:, :, 0], mono, "dup L")
    assert_bit_exact(back[:, :, 1], mono, "dup R")


def test_device_pointer_form_at_full_width():
    import torch

    C, n = 4096, 4096
    e = Engine(C, block=128, max_samples=128)
    st = torch.from_numpy(S.noise(C, 2 * n, seed=9)).cuda()
    mono = torch.empty((C, n), dtype=torch.float32, device="cuda")
    back = torch.empty((C, 2 * n), dtype=torch.float32, device="cuda")
    s = torch.cuda.current_stream()
    e._ck(e._L.dspb_fold_stereo(e._h, st.data_ptr(), mono.data_ptr(), n, MEM_DEVICE, ctypes.c_void_p(s.cuda_stream)))
    e._ck(e._L.dspb_dup_stereo(e._h, mono.data_ptr(), back.data_ptr(), n, MEM_DEVICE, ctypes.c_void_p(s.cuda_stream)))
    torch.cuda.synchronize()
    v = st.view(C, n, 2)
    assert torch.equal(mono, v[:, :, 0] + v[:, :, 1])
    assert torch.equal(back.view(C, n, 2)[:, :, 0], mono) and torch.equal(back.view(C, n, 2)[:, :, 1], mono)
    # the fold composes with the effect path: fold -> engine graph == engine graph on the folded signal (same buffers)


@pytest.mark.parametrize("target,C,calls", [(44100.0, 5, [(4800, 4000), (4800, 4000), (300, 200)]), (96000.0, 3, [(2048, 4000), (1024, 2000)]),
                                            (48000.0, 2, [(512, 500)]), (22050.0, 4, [(4096, 1800), (128, 40)]),
                                            (44100.0, 2, [(100, 400)])])      # the last: output wants more input than given -> zeros
def test_resample_dup_matches_the_restated_dasp_converter(oracle_mod, target, C, calls):
    """dspb_resample_dup_stereo (devices.rs:443-500, 550-556) against the oracle's restatement of dasp's Converter +
    Sinc<[f32; 16]>: bit for bit (the window weights come from the same libm on the host, the kernel does the reference's
    f64 multiply -> f32 round -> f32 add in the reference's order), consumed counts equal, state carried across calls."""
    e = Engine(C, block=128, max_samples=128)
    r = oracle_mod.Resampler(C, target)
    x = S.noise(C, sum(n for n, _ in calls) + 16, seed=3)
    off = 0
    for n_in, n_out in calls:
        blk = x[:, off:off + n_in]
        got, used = e.resample_dup_stereo(blk, n_out, target)
        ref, used_ref = r.process(blk, n_out)
        assert used == used_ref
        assert_bit_exact(got.reshape(C, -1), ref.reshape(C, -1), f"resample to {target} Hz")
        assert np.array_equal(got[:, :, 0], got[:, :, 1])
        off += used
    e.reset_state()   # a fresh converter after reset
    got, _ = e.resample_dup_stereo(x[:, :256], 200, target)
    ref, _ = oracle_mod.Resampler(C, target).process(x[:, :256], 200)
    assert_bit_exact(got.reshape(C, -1), ref.reshape(C, -1), "after reset")
