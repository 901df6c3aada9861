"""FIR node (nodes/fir.rs:179-225): the FFT overlap-save path against the f64 oracle within the float-audio
tolerance, and against the bit-exact direct path; warm-up quirk, Average mode, ragged sizes."""
import numpy as np
import pytest

from dsp_stuff_b200 import signals as S
from tests.test_gpu_parity import make_engine, run_both
from tests.test_oracle_kat import single
from tests.util import assert_audio_close, assert_bit_exact, err_metrics, make_oracle

pytestmark = pytest.mark.gpu
FIR_FFT, FIR_DIRECT = 0, 1


@pytest.mark.parametrize("n_taps,C,n,chunks", [
    (4096, 5, 128 * 160, None),                 # several segments, odd channel count, warm-up inside call 1
    (4096, 4, 128 * 96, [128 * 8, 128 * 88]),   # warm-up spans two calls
    (300, 3, 128 * 70, [128, 128 * 69]),
    (1, 2, 128 * 65, None),                     # default taps [1.0]
    (4097, 2, 128 * 100, None),                 # longest IR the FFT path takes
])
def test_fft_path_vs_oracle(oracle_mod, n_taps, C, n, chunks):
    spec = S.config4(n_taps)
    x = S.noise(C, n)
    got, ref, eng = run_both(oracle_mod, spec, x, chunks=chunks, fir_mode=FIR_FFT)
    rel, dbfs = assert_audio_close(got[0], ref[0], what=f"fir fft N={n_taps}")
    # the first N-1 samples come from the exact warm-up path (fir.rs:192-216) and are bit-identical
    w = min(n, n_taps - 1)
    assert_bit_exact(got[0][:, :w], ref[0][:, :w], "fir warm-up")
    print(f"fir fft N={n_taps}: peak-relative {rel:.2e}, rms {dbfs:.1f} dBFS")


def test_fft_average_mode_and_sweep_input(oracle_mod):
    g = single("fir", mode="Average")
    g.nodes[0].taps = S.reverb_ir(1024)[::-1].copy()
    x = S.sweep(3, 128 * 90)
    got, ref, _ = run_both(oracle_mod, g, x, fir_mode=FIR_FFT)
    # Average mode scales by 1/N: compare relative to that output's own peak
    assert_audio_close(got[0], ref[0], what="fir average (fft)")


def test_fft_matches_direct_at_full_width():
    """BASELINE config 4 width (4096 channels): FFT path against the bit-exact direct path on the device."""
    C, n = 4096, 128 * 40
    spec = S.config4(4096)
    x = S.noise(C, n)
    a = make_engine(spec, C, n, fir_mode=FIR_FFT).process(x)[0]
    sel = np.arange(0, C, 257)
    b = make_engine(spec, len(sel), n, fir_mode=FIR_DIRECT).process(x[sel])[0]
    assert_audio_close(a[sel], b, what="fft vs direct")


def test_target_chain_fft(oracle_mod):
    spec = S.target_chain(4096)
    x = S.noise(6, 128 * 120)
    got, ref, eng = run_both(oracle_mod, spec, x, chunks=[128 * 40, 128 * 80], fir_mode=FIR_FFT)
    assert_audio_close(got[0], ref[0], what="target chain (fft fir)")


def test_long_ir_falls_back_to_exact_path(oracle_mod):
    spec = S.config4(5000)
    x = S.noise(2, 128 * 50)
    got, ref, eng = run_both(oracle_mod, spec, x, fir_mode=FIR_FFT)
    assert_bit_exact(got[0], ref[0], "long IR uses the direct path")


# ---- Toeplitz tensor-core path (csrc/fir_toeplitz.cu): tcgen05 GEMM on hi/lo bf16 operands, f32 accumulate ----
FIR_TOEPLITZ = 2


@pytest.mark.parametrize("n_taps,C,n,chunks", [
    (4096, 5, 128 * 100, None),                 # one partial channel block, warm-up inside the call
    (4096, 300, 128 * 48, [128 * 8, 128 * 40]), # two channel blocks (one partial), state carried across calls
    (300, 3, 128 * 70, [128, 128 * 69]),
    (1, 2, 128 * 33, None),                     # default taps [1.0]: hi + lo must reproduce x to 2^-17
    (4500, 2, 128 * 80, None),                  # longer than the FFT path takes
])
def test_toeplitz_path_vs_oracle(oracle_mod, n_taps, C, n, chunks):
    spec = S.config4(n_taps)
    x = S.noise(C, n)
    got, ref, eng = run_both(oracle_mod, spec, x, chunks=chunks, fir_mode=FIR_TOEPLITZ)
    rel, dbfs = assert_audio_close(got[0], ref[0], what=f"fir toeplitz N={n_taps}")
    w = min(n, n_taps - 1)
    assert_bit_exact(got[0][:, :w], ref[0][:, :w], "fir warm-up")
    print(f"fir toeplitz N={n_taps}: peak-relative {rel:.2e}, rms {dbfs:.1f} dBFS")


def test_toeplitz_matches_fft_at_full_width():
    """BASELINE config 4 (4096 channels): the two throughput paths against each other on the device."""
    C, n = 4096, 128 * 40
    spec = S.config4(4096)
    x = S.noise(C, n)
    a = make_engine(spec, C, n, fir_mode=FIR_FFT).process(x)[0]
    b = make_engine(spec, C, n, fir_mode=FIR_TOEPLITZ).process(x)[0]
    assert_audio_close(b, a, what="toeplitz vs fft")


def test_target_chain_toeplitz(oracle_mod):
    spec = S.target_chain(4096)
    x = S.noise(6, 128 * 120)
    got, ref, eng = run_both(oracle_mod, spec, x, chunks=[128 * 40, 128 * 80], fir_mode=FIR_TOEPLITZ)
    assert_audio_close(got[0], ref[0], what="target chain (toeplitz fir)")


# ---- packed FFT variant (fir_mode 3): two sub-transforms per f32x2 register pair, same tolerance as the FFT path ----
FIR_FFT_PACKED = 3


@pytest.mark.parametrize("n_taps,C,n,chunks", [
    (4096, 5, 128 * 160, None),
    (4097, 4, 128 * 96, [128 * 8, 128 * 88]),
    (300, 3, 128 * 70, [128, 128 * 69]),       # history shorter than the 4096-sample window overlap
    (1, 2, 128 * 65, None),
])
def test_packed_fft_path_vs_oracle(oracle_mod, n_taps, C, n, chunks):
    spec = S.config4(n_taps)
    x = S.noise(C, n)
    got, ref, eng = run_both(oracle_mod, spec, x, chunks=chunks, fir_mode=FIR_FFT_PACKED)
    rel, dbfs = assert_audio_close(got[0], ref[0], what=f"fir packed fft N={n_taps}")
    print(f"fir packed fft N={n_taps}: peak-relative {rel:.2e}, rms {dbfs:.1f} dBFS")
