"""GPU parity: VideoMixer compositing (src/module/video_mixer.rs:70-250) through the C ABI against the
CPU oracle; u8 planes must be bit-exact."""
import numpy as np
import pytest

from mixlab_b200 import workloads as W

pytestmark = pytest.mark.gpu

FADERS = [0.0, 0.25, 0.5, 0.999, 1.0]          # -> f in {0, 63, 127, 254, 255}


def make_frame(ctx, oracle, w, h, seed):
    lay = oracle.frame_layout(w, h)
    data = W.random_bytes(seed, lay.size)
    return ctx.frame(w, h, data), data, lay


def run_mixer(mxl, ctx, params, inputs, ticks=1, t0=0, mod=None):
    """inputs: {channel: [frame or None per tick]}.  Returns (module, [Output, A, B] video lines)."""
    mod = mod or ctx.module(mxl.MOD_VIDEO_MIXER, params)
    ins = []
    for ch in range(4):
        if ch in inputs:
            vl = ctx.video_line(ticks)
            for k, fr in enumerate(inputs[ch]):
                if fr is not None:
                    vl.set(k, fr)
            ins.append(vl)
        else:
            ins.append(None)
    outs = [ctx.video_line(ticks) for _ in range(3)]
    mod.run_tick(t0, ins, outs)
    return mod, outs


@pytest.mark.parametrize("size", [(1920, 1080), (560, 350), (1120, 700), (64, 36), (70, 50), (34, 18)])
def test_crossfade_two_layers(mxl, oracle, ctx48, size):
    w, h = size
    fa, da, lay = make_frame(ctx48, oracle, w, h, 0xA11CE)
    fb, db, _ = make_frame(ctx48, oracle, w, h, 0xB0B)
    flay = mxl.frame_layout(w, h)
    assert (list(flay.stride), list(flay.plane_h), list(flay.offset), flay.size) == (
        list(lay.stride), list(lay.plane_h), list(lay.offset), lay.size)
    for fader in FADERS:
        mod, outs = run_mixer(mxl, ctx48, (0, 1, fader), {0: [fa], 1: [fb]})
        got = outs[0].get(0).download_raw()
        want = oracle.video_crossfade(lay, da, db, oracle.fader_to_u8(fader))
        # bytes the reference never writes (stride padding beyond ceil(w/32)*32, odd last rows) are
        # blank-valued in both
        assert np.array_equal(got, want), (size, fader)
        # pass-through outputs A / B are the input frames themselves (video_mixer.rs:80-90)
        assert np.array_equal(outs[1].get(0).download_raw(), da)
        assert np.array_equal(outs[2].get(0).download_raw(), db)
        mod.destroy()


@pytest.mark.parametrize("missing", ["a", "b", "both_params_none"])
def test_crossfade_missing_layer_aliases_blank(mxl, oracle, ctx48, missing):
    w, h = 1920, 1080
    fa, da, lay = make_frame(ctx48, oracle, w, h, 1)
    if missing == "a":
        params, ins, a, b = (-1, 0, 0.5), {0: [fa]}, None, da
    elif missing == "b":
        params, ins, a, b = (0, -1, 0.5), {0: [fa]}, da, None
    else:
        params, ins, a, b = (-1, -1, 0.5), {0: [fa]}, None, None
    mod, outs = run_mixer(mxl, ctx48, params, ins)
    got = outs[0].get(0).download_raw()
    want = oracle.video_crossfade(lay, a, b, 127)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("fader", [0.0, 1.0])
@pytest.mark.parametrize("missing", ["a", "b"])
def test_end_stop_fader_with_missing_layer(mxl, oracle, ctx48, missing, fader):
    """A layer at weight 0 is not read by the kernel; the result must still be the reference's
    (a*f + b*(255-f)) / 255 with the missing layer aliasing the blank frame (video_mixer.rs:180-188)."""
    w, h = 560, 350
    fa, da, lay = make_frame(ctx48, oracle, w, h, 21)
    params, a, b = ((-1, 0, fader), None, da) if missing == "a" else ((0, -1, fader), da, None)
    for size_case in (0, 1):                                   # flat layout (560 = 17.5 x 32 is padded) and 1080p
        if size_case == 1:
            w, h = 1920, 1080
            fa, da, lay = make_frame(ctx48, oracle, w, h, 22)
            a, b = (None, da) if missing == "a" else (da, None)
        mod, outs = run_mixer(mxl, ctx48, params, {0: [fa]})
        want = oracle.video_crossfade(lay, a, b, oracle.fader_to_u8(fader))
        assert np.array_equal(outs[0].get(0).download_raw(), want), (missing, fader, w)
        mod.destroy()


def test_no_inputs_no_output(mxl, ctx48):
    # video_mixer.rs:113-119: no inputs and no stored pictures -> Output stays None
    mod, outs = run_mixer(mxl, ctx48, (0, 1, 0.5), {})
    assert outs[0].get(0) is None and outs[1].get(0) is None and outs[2].get(0) is None


def test_stored_frame_persists_until_expiry(mxl, oracle, ctx48):
    # video_mixer.rs:92-101,139-143: a frame with duration_hint 1/30 s received at tick 0 is still
    # composited at tick 1 and expires at tick 2 (now >= active_until)
    w, h = 64, 36
    fa, da, lay = make_frame(ctx48, oracle, w, h, 3)
    fb, db, _ = make_frame(ctx48, oracle, w, h, 4)
    mod = ctx48.module(mxl.MOD_VIDEO_MIXER, (0, 1, 0.25))
    ticks = 4
    ina, inb = ctx48.video_line(ticks), ctx48.video_line(ticks)
    ina.set(0, fa, duration=(1, 30))
    for k in range(ticks):
        inb.set(k, fb, duration=(1, 60))
    outs = [ctx48.video_line(ticks) for _ in range(3)]
    mod.run_tick(0, [ina, inb, None, None], outs)
    f = oracle.fader_to_u8(0.25)
    with_a = oracle.video_crossfade(lay, da, db, f)
    without_a = oracle.video_crossfade(lay, None, db, f)
    assert np.array_equal(outs[0].get(0).download_raw(), with_a)
    assert np.array_equal(outs[0].get(1).download_raw(), with_a)
    assert np.array_equal(outs[0].get(2).download_raw(), without_a)
    assert np.array_equal(outs[0].get(3).download_raw(), without_a)
    assert outs[1].get(0) is not None and outs[1].get(1) is None


def test_batched_ticks_equal_single_ticks(mxl, oracle, ctx48):
    w, h = 560, 350
    ticks = 5
    frames_a = [make_frame(ctx48, oracle, w, h, 100 + k) for k in range(ticks)]
    frames_b = [make_frame(ctx48, oracle, w, h, 200 + k) for k in range(ticks)]
    lay = frames_a[0][2]
    mod, outs = run_mixer(mxl, ctx48, (2, 3, 0.7), {2: [f[0] for f in frames_a], 3: [f[0] for f in frames_b]}, ticks=ticks)
    f = oracle.fader_to_u8(0.7)
    for k in range(ticks):
        want = oracle.video_crossfade(lay, frames_a[k][1], frames_b[k][1], f)
        assert np.array_equal(outs[0].get(k).download_raw(), want), k


def test_mixed_sizes_letterbox_geometry(mxl, oracle, ctx48):
    # unify_picture_settings + DynamicScaler (video_mixer.rs:261-297; encode.rs:338-397): a 640x480
    # layer next to a 1280x720 one is letterboxed to 960x720 at x=160 inside a blank 1280x720 frame.
    # The resample itself stands in for third-party swscale (parity unpinned): checked against this
    # repo's own oracle spec.
    fa, da, laya = make_frame(ctx48, oracle, 1280, 720, 7)
    fb, db, layb = make_frame(ctx48, oracle, 640, 480, 8)
    assert mxl.unify_picture_settings(1280, 720, 640, 480) == oracle.unify_picture(1280, 720, 640, 480) == (1280, 720)
    geo = mxl.scale_geometry(640, 480, 1280, 720)
    assert geo == oracle.scale_geometry(640, 480, 1280, 720) == (960, 720, 160, 0)
    scaled = fb.scale(1280, 720)
    got = scaled.download_raw()
    lay = oracle.frame_layout(1280, 720)
    want = oracle.frame_blank(lay)
    for p in range(3):
        sh = 0 if p == 0 else 1
        sw, shh = 640 >> sh, 480 >> sh
        dw, dh = 960 >> sh, 720 >> sh
        src = db[layb.offset[p]:layb.offset[p] + layb.stride[p] * layb.plane_h[p]]
        dst = oracle.bicubic_plane(src, sw, shh, layb.stride[p], dw, dh, dw)
        plane = want[lay.offset[p]:lay.offset[p] + lay.stride[p] * lay.plane_h[p]].reshape(lay.plane_h[p], lay.stride[p])
        plane[0:dh, (160 >> sh):(160 >> sh) + dw] = dst.reshape(dh, dw)
    assert np.array_equal(got, want)
    # identity when sizes match (encode.rs:342-345): the very same frame comes back
    same = fa.scale(1280, 720)
    assert same.h == fa.h
    # and the mixer composites the letterboxed layer
    mod, outs = run_mixer(mxl, ctx48, (0, 1, 0.5), {0: [fa], 1: [fb]})
    out = outs[0].get(0)
    assert (out.layout.width, out.layout.height) == (1280, 720)
    assert np.array_equal(out.download_raw(), oracle.video_crossfade(lay, da, want, 127))


def oracle_scale(oracle, data, lay_in, out_w, out_h):
    return oracle.letterbox_scale(data, lay_in, out_w, out_h)


@pytest.mark.parametrize("src,dst", [((640, 480), (1280, 720)), ((1920, 1080), (1280, 720)), ((1280, 720), (1920, 1080)),
                                     ((3840, 2160), (1920, 1080)), ((70, 50), (560, 350)), ((560, 350), (70, 50)),
                                     ((1920, 1080), (480, 270)), ((34, 18), (36, 20)), ((1000, 1000), (1920, 1080)),
                                     ((1920, 1080), (3840, 2160)), ((2, 2), (64, 64)), ((1366, 768), (1920, 1080)),
                                     ((3840, 2160), (320, 180)), ((3840, 2160), (96, 54)), ((4096, 64), (64, 64))])
def test_tiled_scaler_matches_two_pass_definition(mxl, oracle, ctx48, src, dst):
    """Every tile of the fused scaler (clamped aprons, ragged last tiles, up- and down-scaling, letterbox
    bars on either axis) against the oracle's plain two-pass definition.  UNPINNED arithmetic (stands in
    for swscale), pinned geometry."""
    fr, data, lay_in = make_frame(ctx48, oracle, src[0], src[1], src[0] * 31 + dst[1])
    got = fr.scale(dst[0], dst[1])
    assert (got.layout.width, got.layout.height) == dst
    assert np.array_equal(got.download_raw(), oracle_scale(oracle, data, lay_in, dst[0], dst[1]))


def test_mixer_scales_all_ticks_of_a_call_in_one_launch(mxl, oracle, ctx48):
    """Six ticks, layer B smaller than layer A on every tick: one blank fill per letterboxed frame, ONE
    scaler launch for all six frames, one crossfade launch (video_mixer.rs:122-148, encode.rs:338-397)."""
    ticks = 6
    fa = [make_frame(ctx48, oracle, 1280, 720, 600 + k) for k in range(ticks)]
    fb = [make_frame(ctx48, oracle, 640, 480, 700 + k) for k in range(ticks)]
    before = ctx48.launch_count
    mod, outs = run_mixer(mxl, ctx48, (0, 1, 0.3), {0: [f[0] for f in fa], 1: [f[0] for f in fb]}, ticks=ticks)
    assert ctx48.launch_count - before == ticks + 2
    lay = oracle.frame_layout(1280, 720)
    f8 = oracle.fader_to_u8(0.3)
    for k in range(ticks):
        scaled = oracle_scale(oracle, fb[k][1], fb[k][2], 1280, 720)
        assert np.array_equal(outs[0].get(k).download_raw(), oracle.video_crossfade(lay, fa[k][1], scaled, f8)), k


def test_target_growing_inside_a_call_rescales_finished_pixels(mxl, oracle, ctx48):
    """Channel A delivers one 320x180 frame at tick 0 that stays stored (long duration); channel B delivers 640x360
    frames, then 1280x720 from tick 3 on.  At tick 0 A is scaled up to the unified 640x360 target and stored -- in a
    batched call by the deferred scaler launch.  At tick 3 the target grows and A's STORED (scaled) frame is scaled
    again, immediately (Channel::rescale, video_mixer.rs:262-274): it must read finished pixels, not a frame whose
    scaler launch is still pending.  Six ticks in one call must equal six one-tick calls, and both the oracle."""
    ticks = 6
    fa = make_frame(ctx48, oracle, 320, 180, 4242)
    fb = [make_frame(ctx48, oracle, 640, 360, 5000 + k) for k in range(3)] + [make_frame(ctx48, oracle, 1280, 720, 6000 + k) for k in range(3)]
    f8 = oracle.fader_to_u8(0.4)

    def feed(n_calls):
        mod = ctx48.module(mxl.MOD_VIDEO_MIXER, (0, 1, 0.4))
        got = []
        per = ticks // n_calls
        for c in range(n_calls):
            la, lb = ctx48.video_line(per), ctx48.video_line(per)
            for k in range(per):
                tick = c * per + k
                if tick == 0:
                    la.set(k, fa[0], duration=(10, 1))            # stays stored for the whole test
                lb.set(k, fb[tick][0], duration=(1, 60))
            outs = [ctx48.video_line(per) for _ in range(3)]
            mod.run_tick(c * per * 800, [la, lb, None, None], outs)
            for k in range(per):
                o = outs[0].get(k)
                got.append(((o.layout.width, o.layout.height), o.download_raw()))
        mod.destroy()
        return got

    batched, single = feed(1), feed(6)
    a_mid = oracle_scale(oracle, fa[1], fa[2], 640, 360)
    a_big = oracle_scale(oracle, a_mid, oracle.frame_layout(640, 360), 1280, 720)      # the stored frame scaled again
    for k in range(ticks):
        tw, th = (640, 360) if k < 3 else (1280, 720)
        lay = oracle.frame_layout(tw, th)
        want = oracle.video_crossfade(lay, a_mid if k < 3 else a_big, fb[k][1], f8)
        assert single[k][0] == (tw, th) and batched[k][0] == (tw, th), k
        assert np.array_equal(single[k][1], want), ("one tick per call", k)
        assert np.array_equal(batched[k][1], want), ("six ticks in one call", k)


def test_batched_scaler_one_launch(mxl, oracle, ctx48):
    frames = [make_frame(ctx48, oracle, 640, 360, 900 + k) for k in range(6)]
    before = ctx48.launch_count
    out = ctx48.frames_scale([f[0] for f in frames], 1280, 720)        # same aspect: no bars, no blank fill
    assert ctx48.launch_count - before == 1
    for k, fr in enumerate(out):
        assert np.array_equal(fr.download_raw(), oracle_scale(oracle, frames[k][1], frames[k][2], 1280, 720)), k
    same = ctx48.frames_scale([f[0] for f in frames], 640, 360)        # identity: the very same frames (encode.rs:342-345)
    assert [f.h for f in same] == [f[0].h for f in frames]


@pytest.mark.parametrize("size", [(1920, 1080), (560, 350), (70, 50), (34, 18)])
def test_compose_rgba_equals_crossfade_then_convert(mxl, oracle, ctx48, size):
    """The one-pass compositor output (blend + colour conversion) is bit-identical to VideoMixer's
    crossfade (pinned, video_mixer.rs:211-235) followed by the self-specified yuv420p -> RGBA."""
    w, h = size
    n = 5
    la = [make_frame(ctx48, oracle, w, h, 300 + k) for k in range(n)]
    lb = [make_frame(ctx48, oracle, w, h, 400 + k) for k in range(n)]
    lay = la[0][2]
    pics = ctx48.rgba(w, h, n + 1)
    for fader in (0.0, 0.25, 0.5, 0.999, 1.0):
        before = ctx48.launch_count
        ctx48.compose_rgba([f[0] for f in la], [f[0] for f in lb], fader, pics, first=1)
        assert ctx48.launch_count - before == 1
        got = pics.download(1, n)
        f8 = oracle.fader_to_u8(fader)
        for k in range(n):
            want = oracle.yuv420p_to_rgba(lay, oracle.video_crossfade(lay, la[k][1], lb[k][1], f8))
            assert np.array_equal(got[k], want), (fader, k)
    # a missing layer is blank (video_mixer.rs:180-188), on either side
    ctx48.compose_rgba([la[0][0], None], [None, lb[1][0]], 0.25, pics, first=0)
    got = pics.download(0, 2)
    f8 = oracle.fader_to_u8(0.25)
    assert np.array_equal(got[0], oracle.yuv420p_to_rgba(lay, oracle.video_crossfade(lay, la[0][1], None, f8)))
    assert np.array_equal(got[1], oracle.yuv420p_to_rgba(lay, oracle.video_crossfade(lay, None, lb[1][1], f8)))
    # plain conversion of a batch = compose with a single layer
    ctx48.frames_to_rgba([f[0] for f in la], pics, first=0)
    got = pics.download(0, n)
    for k in range(n):
        assert np.array_equal(got[k], oracle.yuv420p_to_rgba(lay, la[k][1]))
    pics.free()


def test_frame_batch_moves_as_one_copy(mxl, oracle, ctx48):
    """mxl_frames_alloc_batch: adjacent frames, uploaded / downloaded as one copy per run, usable like any frame,
    freed with the last of them whatever the release order."""
    import ctypes as C
    w, h, n = 1920, 1080, 6
    lay = oracle.frame_layout(w, h)
    frames = ctx48.frames_batch(w, h, n)
    ptrs = [mxl.lib().mxl_frame_device_ptr(f.h) for f in frames]
    assert all(ptrs[i + 1] - ptrs[i] == lay.size for i in range(n - 1))
    host = mxl.PinnedBuffer(n * lay.size)
    host.array[:] = W.random_bytes(1234, n * lay.size)
    arr = (C.c_void_p * n)(*[f.h for f in frames])
    before = ctx48.h2d_bytes
    mxl.check(mxl.lib().mxl_frames_upload_raw_async(arr, n, host.ptr, lay.size))
    ctx48.synchronize()
    assert ctx48.h2d_bytes - before == n * lay.size
    for k, f in enumerate(frames):
        assert np.array_equal(f.download_raw(), host.array[k * lay.size:(k + 1) * lay.size]), k
    # a mixed list (batch frames interleaved with a pool frame) still lands frame by frame
    other = ctx48.frame(w, h, blank=True)
    mixed = [frames[3], other, frames[0], frames[1]]
    marr = (C.c_void_p * 4)(*[f.h for f in mixed])
    back = mxl.PinnedBuffer(4 * lay.size)
    mxl.check(mxl.lib().mxl_frames_download_raw_async(marr, 4, back.ptr, lay.size))
    ctx48.synchronize()
    got = back.array.reshape(4, lay.size)
    assert np.array_equal(got[0], frames[3].download_raw()) and np.array_equal(got[1], oracle.frame_blank(lay))
    assert np.array_equal(got[2], frames[0].download_raw()) and np.array_equal(got[3], frames[1].download_raw())
    # the mixer takes them like any frame
    mod, outs = run_mixer(mxl, ctx48, (0, 1, 0.5), {0: frames[:3], 1: frames[3:]}, ticks=3)
    for k in range(3):
        want = oracle.video_crossfade(lay, host.array[k * lay.size:(k + 1) * lay.size], host.array[(k + 3) * lay.size:(k + 4) * lay.size], 127)
        assert np.array_equal(outs[0].get(k).download_raw(), want), k
    for f in (frames[2], frames[5], frames[0], frames[4], frames[1], frames[3], other):
        f.release()
    host.free(); back.free()


@pytest.mark.parametrize("size", [(1920, 1080), (560, 350), (70, 50), (34, 18), (35, 19), (9, 3), (2, 2)])
def test_rgba_to_yuv_self_specified(mxl, oracle, ctx48, size):
    """UNPINNED (north_star's "YUV<->RGB", no reference counterpart): RGBA8 -> yuv420p against the oracle's integer
    definition, byte for byte incl. odd pictures (cut chroma blocks) and untouched stride padding; n pictures, one launch."""
    w, h = size
    n = 3
    lay = oracle.frame_layout(w, h)
    rgba = W.random_bytes(w * 131 + h, n * w * h * 4)
    pics = ctx48.rgba(w, h, n)
    pics.upload(rgba)
    pad = [W.random_bytes(900 + k, lay.size) for k in range(n)]          # what the frames held before: padding must survive
    frames = [ctx48.frame(w, h, pad[k]) for k in range(n)]
    before = ctx48.launch_count
    pics.to_frames(frames)
    assert ctx48.launch_count - before == 1
    for k in range(n):
        want = oracle.rgba_to_yuv420p(lay, rgba[k * w * h * 4:(k + 1) * w * h * 4], into=pad[k])
        assert np.array_equal(frames[k].download_raw(), want), k
    # and through the compositor: a picture converted to yuv420p enters VideoMixer like any frame
    mod, outs = run_mixer(mxl, ctx48, (0, -1, 1.0), {0: [frames[0]]})
    assert np.array_equal(outs[0].get(0).download_raw()[:lay.offset[1]].reshape(lay.plane_h[0], lay.stride[0])[:, :w],
                          oracle.rgba_to_yuv420p(lay, rgba[:w * h * 4]) [:lay.offset[1]].reshape(lay.plane_h[0], lay.stride[0])[:, :w])
    pics.free()


@pytest.mark.parametrize("size", [(1920, 1080), (70, 50)])
def test_rgb_yuv_rgb_round_trip(mxl, oracle, ctx48, size):
    """The two self-specified conversions are inverse to each other up to quantisation: a picture whose 2x2 blocks are
    flat (so the chroma mean loses nothing) goes RGBA8 -> yuv420p -> RGBA8 through both kernels and comes back within
    3 levels per channel (0.5 of rounding in each of Y, U, V times the BT.601 gains 1.164 / 2.018, plus the final rounding),
    alpha 255.  Colours are kept inside the limited-range gamut so that no clip is involved."""
    w, h = size
    rng = np.random.default_rng(w * 7 + h)
    blocks = rng.integers(40, 216, size=(h // 2, w // 2, 3), dtype=np.uint8)
    rgb = np.repeat(np.repeat(blocks, 2, axis=0), 2, axis=1)
    rgba = np.concatenate([rgb, np.full((h, w, 1), 255, np.uint8)], axis=2).reshape(-1)
    pics = ctx48.rgba(w, h, 1)
    pics.upload(rgba)
    fr = ctx48.frame(w, h, blank=True)
    pics.to_frames([fr])
    back = fr.to_rgba().reshape(h, w, 4)
    assert np.all(back[:, :, 3] == 255)
    err = np.abs(back[:, :, :3].astype(np.int32) - rgb.astype(np.int32))
    assert err.max() <= 3, err.max()
    # the device round trip is the oracle's round trip, byte for byte
    lay = oracle.frame_layout(w, h)
    want = oracle.yuv420p_to_rgba(lay, oracle.rgba_to_yuv420p(lay, rgba))
    assert np.array_equal(back.reshape(-1), want)
    pics.free()


def test_yuv_to_rgba_self_specified(mxl, oracle, ctx48):
    # UNPINNED: the reference never converts colour (video_mixer.rs:282-283); spec = oracle header
    for (w, h) in [(1920, 1080), (70, 50), (34, 18)]:
        fr, data, lay = make_frame(ctx48, oracle, w, h, 0xC0FFEE + w)
        assert np.array_equal(fr.to_rgba(), oracle.yuv420p_to_rgba(lay, data))


def test_blank_frame(mxl, oracle, ctx48):
    for (w, h) in [(1920, 1080), (560, 350), (34, 18)]:
        fr = ctx48.frame(w, h, blank=True)
        assert np.array_equal(fr.download_raw(), oracle.frame_blank(oracle.frame_layout(w, h)))


def test_plane_upload_download_with_caller_strides(mxl, oracle, ctx48):
    w, h = 70, 50
    fr = ctx48.frame(w, h, blank=True)
    planes = [W.random_bytes(1, 80 * 50), W.random_bytes(2, 40 * 25), W.random_bytes(3, 40 * 25)]
    fr.upload_planes(planes, (80, 40, 40))
    back = fr.download_planes((80, 40, 40))
    for p, (pw, ph, st) in enumerate([(70, 50, 80), (35, 25, 40), (35, 25, 40)]):
        a = planes[p].reshape(ph, st)[:, :pw]
        b = back[p].reshape(ph, st)[:, :pw]
        assert np.array_equal(a, b)
