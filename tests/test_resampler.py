"""Audio sample-rate converter (north_star "resample"): NEW and self-specified -- the reference has only the TODO
(Icecast ingest drops every stream that is not at the engine's rate, src/icecast/mod.rs:94-97).  PARITY UNPINNED: the
oracle (oracle/mixlab_oracle.c orc_resampler_*) is the definition; the CPU tests pin the definition's properties, the GPU
tests the kernel against it bit for bit."""
import numpy as np
import pytest

from mixlab_b200 import workloads as W


def i16(seed, n):
    return np.ascontiguousarray(W.random_bytes(seed, 2 * n)).view(np.int16).copy()


def test_definition_ratio_latency_and_count(oracle):
    r = oracle.Resampler(44100, 48000, 1)
    assert (r.L, r.M) == (160, 147)
    total_in = total_out = 0
    for n in (1, 15, 1, 100, 735, 4410):
        total_out += len(r.push(np.zeros(n, np.float32)))
        total_in += n
        want = 0 if total_in <= 16 else -(-(total_in - 16) * 160 // 147)          # ceil((N - 16) L / M)
        assert total_out == want, (total_in, total_out, want)
    r = oracle.Resampler(48000, 44100, 2)
    assert (r.L, r.M) == (147, 160)


def test_definition_dc_gain_and_impulse_response(oracle):
    r = oracle.Resampler(44100, 48000, 1)
    y = r.push(np.ones(3000, np.float32))
    assert np.max(np.abs(y[64:] - 1.0)) <= 1e-7                 # every phase row sums to 1
    # an impulse at input frame 200 comes out as the coefficient table: y[m] = c[phase(m)][k] with n0(m) - 15 + k = 200
    r = oracle.Resampler(44100, 48000, 1)
    x = np.zeros(2000, np.float32)
    x[200] = 1.0
    y = r.push(x)
    seen = 0
    for m in range(len(y)):
        n0, phase = (m * 147) // 160, (m * 147) % 160
        k = 200 - n0 + 15
        want = np.float32(r.coef(phase, k)) if 0 <= k < 32 else np.float32(0.0)
        assert y[m] == want, (m, k)
        seen += 0 <= k < 32
    assert seen >= 32 * 160 // 147


def test_definition_is_a_function_of_the_stream_not_of_the_calls(oracle):
    x = W.uniform_pm1(77, 2 * 9000)
    whole = oracle.Resampler(44100, 48000, 2).push(x)
    r = oracle.Resampler(44100, 48000, 2)
    rng = np.random.default_rng(5)
    parts, pos = [], 0
    while pos < 9000:
        n = int(min(9000 - pos, rng.integers(1, 1500)))
        parts.append(r.push(x[2 * pos:2 * (pos + n)]))
        pos += n
    assert np.array_equal(np.concatenate(parts).view(np.uint32), whole.view(np.uint32))


@pytest.mark.parametrize("rates", [(44100, 48000), (48000, 44100), (32000, 48000)])
def test_definition_passes_a_tone(oracle, rates):
    """a 1 kHz sine arrives as a 1 kHz sine at the new rate: error below -100 dB (f32 rounding is at -140 dB)."""
    fin, fout = rates
    n = fin // 2
    x = np.sin(2 * np.pi * 1000.0 * np.arange(n) / fin).astype(np.float32)
    y = oracle.Resampler(fin, fout, 1).push(x)
    ref = np.sin(2 * np.pi * 1000.0 * np.arange(len(y)) / fout)
    err = y[256:].astype(np.float64) - ref[256:]
    assert 20 * np.log10(np.sqrt(np.mean(err ** 2)) / np.sqrt(0.5)) < -100.0


@pytest.mark.gpu
@pytest.mark.parametrize("rates,channels", [((44100, 48000), 2), ((48000, 44100), 1), ((22050, 48000), 2), ((48000, 16000), 2)])
def test_device_resampler_i16_pushes_equal_the_definition(mxl, oracle, ctx48, rates, channels):
    """What a receiver does with a 44.1 kHz source: decoded i16 frames of irregular size in, f32 at the engine's rate out.
    Every push bit-exact against the definition, output counts as predicted."""
    fin, fout = rates
    rs = ctx48.resampler(fin, fout, channels)
    orc = oracle.Resampler(fin, fout, channels)
    rng = np.random.default_rng(fin + channels)
    total = 0
    for i, n in enumerate([1, 7, 16, 1024, 1024, 333, 4096, 20000, 2, 1152]):
        pcm = i16(1000 + i, n * channels)
        predicted = rs.output_frames(n)
        got = rs.push_i16(pcm)
        want = orc.push(pcm)
        assert got.size == want.size == predicted * channels, (i, got.size, want.size, predicted)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), i
        total += got.size
    assert total > 0
    rs.reset()
    orc2 = oracle.Resampler(fin, fout, channels)
    pcm = i16(5, 5000 * channels)
    assert np.array_equal(rs.push_i16(pcm).view(np.uint32), orc2.push(pcm).view(np.uint32))
    rs.close()


@pytest.mark.gpu
def test_device_resampler_from_a_line_and_in_one_launch(mxl, oracle, ctx48):
    x = W.uniform_pm1(31, 2 * 44100)
    rs = ctx48.resampler(44100, 48000, 2)
    line = ctx48.stereo(x)
    before = ctx48.launch_count
    got = rs.push_line(line)
    assert ctx48.launch_count - before == 2                     # the converter and its 32-frame history hand-over
    want = oracle.Resampler(44100, 48000, 2).push(x)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert got.size // 2 == -(-(44100 - 16) * 160 // 147)
    rs.close()
    line.free()
