// video_kernels.cu -- VideoMixer compositing kernels for sm_100a (src/module/video_mixer.rs:150-239).
//
// Crossfade: out = (a*f + b*(255-f)) / 255 per byte, u16 lanes, truncating (fade_line, 211-235).
// HBM-bound integer work: 16-byte (uint4) accesses, two per thread per layer for memory-level
// parallelism, and SIMD-in-register arithmetic -- two bytes per 32-bit multiply:
//     even/odd bytes are split into 16-bit lanes (lane value <= 255), x = lane_a*f + lane_b*(255-f)
//     stays <= 65025 per lane so nothing carries across lanes, and the exact truncating /255 is
//     (x + 1 + (x >> 8)) >> 8 (identity checked exhaustively for 0..65025 in tests/test_host_logic.py),
//     whose intermediate (<= 65280) also fits the lane.
// The reference first memsets a blank frame (AvFrame::blank, frame.rs:94-135) and lets a missing
// layer alias it; here a missing layer is synthesised in registers (Y=0x00, U=V=0x80), which
// removes the memset pass and the re-read of the blank plane.
#include "kernels.h"

namespace mxl {
namespace k {

namespace {

constexpr int kVidThreads = 256;

__device__ __forceinline__ uint4 ldg16(const uint8_t* p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

__device__ __forceinline__ uint32_t fade4(uint32_t a, uint32_t b, uint32_t f, uint32_t g)
{
    const uint32_t ae = a & 0x00FF00FFu, ao = (a >> 8) & 0x00FF00FFu;
    const uint32_t be = b & 0x00FF00FFu, bo = (b >> 8) & 0x00FF00FFu;
    const uint32_t xe = ae * f + be * g;
    const uint32_t xo = ao * f + bo * g;
    const uint32_t ye = xe + 0x00010001u + ((xe >> 8) & 0x00FF00FFu);
    const uint32_t yo = xo + 0x00010001u + ((xo >> 8) & 0x00FF00FFu);
    return ((ye >> 8) & 0x00FF00FFu) | (yo & 0xFF00FF00u);
}

__device__ __forceinline__ uint4 fade16(uint4 a, uint4 b, uint32_t f, uint32_t g)
{
    return make_uint4(fade4(a.x, b.x, f, g), fade4(a.y, b.y, f, g), fade4(a.z, b.z, f, g), fade4(a.w, b.w, f, g));
}

// Flat path: every plane has stride == processed row width and rows == plane height, so the
// processed bytes of a frame are one contiguous range [0, size).
constexpr int kFadeUnroll = 2;     // 4 measured slower: 0.79 vs 0.96 of the copy peak (64 frames)

__device__ __forceinline__ void crossfade_flat_body(const FadeJob job, uint64_t n16, uint64_t chroma16)
{
    const uint32_t f = job.fade, g = 255u - job.fade;
    // the kFadeUnroll vectors of a thread are kVidThreads apart: each warp instruction is 512 contiguous bytes
    const uint64_t v0 = (uint64_t)blockIdx.x * (kVidThreads * kFadeUnroll) + threadIdx.x;
    uint4 a[kFadeUnroll], b[kFadeUnroll];
#pragma unroll
    for (int u = 0; u < kFadeUnroll; u++) {
        const uint64_t v = v0 + (uint64_t)u * kVidThreads;
        if (v < n16) {
            const uint32_t blank = v >= chroma16 ? 0x80808080u : 0u;
            // (the host passes a layer whose weight is 0 -- fader at an end stop -- as missing: it is not read)
            a[u] = job.a ? ldg16(job.a + v * 16) : make_uint4(blank, blank, blank, blank);
            b[u] = job.b ? ldg16(job.b + v * 16) : make_uint4(blank, blank, blank, blank);
        }
    }
#pragma unroll
    for (int u = 0; u < kFadeUnroll; u++) {
        const uint64_t v = v0 + (uint64_t)u * kVidThreads;
        if (v < n16) *reinterpret_cast<uint4*>(job.out + v * 16) = fade16(a[u], b[u], f, g);
    }
}

__global__ void __launch_bounds__(kVidThreads) crossfade_flat_kernel(const FadeJob* __restrict__ jobs, uint64_t n16, uint64_t chroma16)
{
    crossfade_flat_body(jobs[blockIdx.y], n16, chroma16);
}

// A call of a few ticks (the 60 Hz engine thread: one) carries its job table in the kernel parameters: no
// table upload precedes the launch.
__global__ void __launch_bounds__(kVidThreads) crossfade_flat_inline_kernel(const __grid_constant__ FadeJobsInline jobs, uint64_t n16, uint64_t chroma16)
{
    crossfade_flat_body(jobs.job[blockIdx.y], n16, chroma16);
}

// General path: one plane per launch, rows may carry untouched stride padding.
__device__ __forceinline__ void crossfade_plane_body(const FadeJob job, uint64_t plane_offset, uint32_t stride, uint32_t vec_per_row, uint32_t blank)
{
    const uint32_t col = blockIdx.x * kVidThreads + threadIdx.x;
    if (col >= vec_per_row) return;
    const uint64_t off = plane_offset + (uint64_t)blockIdx.y * stride + (uint64_t)col * 16;
    const uint32_t f = job.fade, g = 255u - job.fade;
    const uint4 a = job.a ? ldg16(job.a + off) : make_uint4(blank, blank, blank, blank);
    const uint4 b = job.b ? ldg16(job.b + off) : make_uint4(blank, blank, blank, blank);
    *reinterpret_cast<uint4*>(job.out + off) = fade16(a, b, f, g);
}

__global__ void __launch_bounds__(kVidThreads) crossfade_plane_kernel(const FadeJob* __restrict__ jobs, uint64_t plane_offset,
                                                                      uint32_t stride, uint32_t vec_per_row, uint32_t blank)
{
    crossfade_plane_body(jobs[blockIdx.z], plane_offset, stride, vec_per_row, blank);
}

__global__ void __launch_bounds__(kVidThreads) crossfade_plane_inline_kernel(const __grid_constant__ FadeJobsInline jobs, uint64_t plane_offset,
                                                                             uint32_t stride, uint32_t vec_per_row, uint32_t blank)
{
    crossfade_plane_body(jobs.job[blockIdx.z], plane_offset, stride, vec_per_row, blank);
}

__global__ void __launch_bounds__(kVidThreads) blank_kernel(uint4* dst, uint64_t n16, uint64_t chroma16)
{
    uint64_t i = (uint64_t)blockIdx.x * kVidThreads + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * kVidThreads;
    for (; i < n16; i += stride) {
        const uint32_t w = i >= chroma16 ? 0x80808080u : 0u;
        dst[i] = make_uint4(w, w, w, w);
    }
}

// yuv420p -> RGBA8, BT.601 limited range, integer (self-specified; see DESIGN.md "unpinned").
__device__ __forceinline__ uint32_t clip8(int v) { return (uint32_t)min(max(v, 0), 255); }

__device__ __forceinline__ uint32_t yuv_px(int y, int u, int v)
{
    const int c = y - 16, d = u - 128, e = v - 128;
    const uint32_t r = clip8((298 * c + 409 * e + 128) >> 8);
    const uint32_t g = clip8((298 * c - 100 * d - 208 * e + 128) >> 8);
    const uint32_t b = clip8((298 * c + 516 * d + 128) >> 8);
    return r | (g << 8) | (b << 16) | 0xFF000000u;
}

// The same conversion arranged for throughput (the compositor is integer-ALU-bound, not HBM-bound, when every
// pixel runs yuv_px): the chroma terms of a 2x2 block are formed once, each channel is one multiply-add
// 298*y + term, the clip works on the unshifted sum (clamp to [0, 65535], then byte 1 is the clipped
// (sum >> 8): a negative sum gives 0, a sum >= 65536 gives 255), and bytes are gathered with PRMT.
struct ChromaTerms { int r, g, b; };
__device__ __forceinline__ uint32_t byte_of(uint32_t w, int i) { return __byte_perm(w, 0u, 0x4440u | (uint32_t)i); }
__device__ __forceinline__ ChromaTerms chroma_terms(uint32_t u, uint32_t v)
{
    // 128 (rounding) - 298*16 folded in:  409 e + 128 - 4768,  -100 d - 208 e + 128 - 4768,  516 d + 128 - 4768
    ChromaTerms t;
    t.r = 409 * (int)v - 56992;
    t.g = -100 * (int)u - 208 * (int)v + 34784;
    t.b = 516 * (int)u - 70688;
    return t;
}
__device__ __forceinline__ uint32_t yuv_px_terms(uint32_t y, const ChromaTerms& t)
{
    const int r = __vimin_s32_relu(298 * (int)y + t.r, 65535);
    const int g = __vimin_s32_relu(298 * (int)y + t.g, 65535);
    const int b = __vimin_s32_relu(298 * (int)y + t.b, 65535);
    const uint32_t rg = __byte_perm((uint32_t)r, (uint32_t)g, 0x4451u);      // byte1(r), byte1(g), 0, 0
    return __byte_perm(rg, (uint32_t)b, 0x4510u) | 0xFF000000u;              // r, g, byte1(b), alpha
}

// ------------------------------------------------------------------------------------------------
// Tiled letterbox scaler (DynamicScaler::scale, src/video/encode.rs:338-397; the arithmetic is this
// repository's stand-in for swscale's SWS_BICUBIC -- DESIGN.md "unpinned"): all three planes of a
// batch of frames in ONE launch.  A CTA owns a 128x64 tile of output pixels of one plane (128x32, 128x8 or 128x2 when a
// strong down-scale would make the source rectangle of a taller tile outgrow shared memory):
//   1. the source rectangle the tile's taps touch is staged global -> shared with 16-byte cp.async; tiles at the
//      left / right edge of the plane then replicate the edge column outwards, so that the four taps of a pixel are
//      always four consecutive staged bytes (the taps clamp);
//   2. horizontal 4-tap pass shared -> shared into a u8 intermediate (rounded and clipped exactly as the
//      two-pass definition does): two aligned word loads + one byte permute per pixel instead of four byte
//      gathers;
//   3. vertical 4-tap pass shared -> global.
// Every source byte is read from HBM once per tile that needs it (neighbouring tiles share a 3-pixel
// apron through L2); the intermediate never leaves the SM.
// ------------------------------------------------------------------------------------------------
constexpr int kScaleTW = 128;     // tile width; the tile height is a template parameter (128, 64, 32, 8 or 2 output rows)

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src)
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gmem_src) : "memory");
}

// Shared-memory accesses through 32-bit shared-window addresses held in registers: with generic pointers to static
// __shared__ arrays the compiler re-derived the window base (S2UR CgaCtaId + uniform adds) inside the tap loops.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint2 lds64(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ uint4 lds128(uint32_t a) { uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void cp_async_16s(uint32_t smem_dst, const void* gmem_src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_dst), "l"(gmem_src) : "memory");
}

// Four taps in two instructions: weights as two s16 pairs, the four u8 samples in one word (dp2a.lo takes the
// two low bytes, dp2a.hi the two high ones); then the two-pass definition's rounding and clip.
__device__ __forceinline__ int pack_s16x2(short lo, short hi) { return (int)(((uint32_t)(uint16_t)hi << 16) | (uint16_t)lo); }
__device__ __forceinline__ int dp2a_lo(int c, uint32_t px, int acc)
{
    int r;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(c), "r"(px), "r"(acc));
    return r;
}
__device__ __forceinline__ int dp2a_hi(int c, uint32_t px, int acc)
{
    int r;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(c), "r"(px), "r"(acc));
    return r;
}
// the two-pass definition's rounding (+8192 before, >> 14 after) for the four taps of one pixel held in one word
__device__ __forceinline__ int tap4(uint32_t px, int c01, int c23) { return dp2a_hi(c23, px, dp2a_lo(c01, px, 8192)) >> 14; }
// four pixel sums -> four bytes, each clipped to [0, 255]: two saturating pack instructions (I2IP) instead of four min/max
// and three byte permutes
__device__ __forceinline__ uint32_t pack4_sat_u8(int v0, int v1, int v2, int v3)
{
    uint32_t hi, all;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(v3), "r"(v2), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(all) : "r"(v1), "r"(v0), "r"(hi));
    return all;                                            // bytes v0, v1, v2, v3
}

template <int kScaleTH>
__global__ void __launch_bounds__(kVidThreads) scale_tiled_kernel(const __grid_constant__ ScaleLaunch L)
{
    extern __shared__ __align__(16) uint8_t sc_smem[];
    // per output column: {byte offset of the aligned word holding the first tap, PRMT selector of the four taps,
    // weights 0|1, weights 2|3}; per output row: the byte offsets of the four tap rows in `mid`, and the weights
    __shared__ uint4 s_x[kScaleTW];
    __shared__ uint4 s_yr[kScaleTH];
    __shared__ uint2 s_yc[kScaleTH];
    const ScaleJob job = L.jobs[blockIdx.y];
    // the plane's geometry into registers through statically addressed parameter reads (a run-time index into
    // the parameter block turns every field access into a slow indexed constant load)
    ScalePlane P = L.pl[0];
    if (blockIdx.x >= L.pl[2].tile_base) P = L.pl[2];
    else if (blockIdx.x >= L.pl[1].tile_base) P = L.pl[1];
    const uint32_t t = blockIdx.x - P.tile_base;
    const uint32_t tyi = P.tiles_x == 1 ? t : __umulhi(t, P.tiles_x_magic);   // t / tiles_x (magic = ceil(2^32 / tiles_x))
    const uint32_t x0 = (t - tyi * P.tiles_x) * kScaleTW, y0 = tyi * kScaleTH;
    const uint32_t x1 = min(x0 + kScaleTW, P.dst_w) - 1, y1 = min(y0 + kScaleTH, P.dst_h) - 1;
    const int sw = (int)P.src_w, sh = (int)P.src_h;
    // source rectangle touched by the tile's taps (first taps are monotonic in the output index); the first taps of
    // the tile's corner columns / rows come from the tap tables (four warp-uniform loads, L2 hits after the first
    // tile of a geometry).  Columns are kept UNclamped: the staged rows carry replicated edge pixels left of column
    // 0 and right of column sw-1, so the four taps of a pixel are always four consecutive staged bytes.
    const int px_lo = __ldg(P.xpos + x0), px_hi = __ldg(P.xpos + x1) + 3;       // px_lo >= -2, px_hi <= sw + 1
    const int gx0 = px_lo & ~15;                                                 // floor to 16 (-16 for a negative first tap)
    const int width = (px_hi - gx0 + 5 + 15) & ~15;                              // the second word of the last tap set included
    const int ry_lo = min(max(__ldg(P.ypos + y0), 0), sh - 1);
    const int ry_hi = min(max(__ldg(P.ypos + y1) + 3, 0), sh - 1);
    const int nch = width >> 4, rows = ry_hi - ry_lo + 1;
    const int pitch = (int)L.region_pitch;                       // bytes per staged row (host bound, multiple of 16)
    uint8_t* region = sc_smem;                                   // [rows][pitch]
    uint8_t* mid = sc_smem + (size_t)L.region_rows * pitch;      // [rows][kScaleTW]
    const uint8_t* src = job.src + P.src_off;
    // 16 threads per staged row, 16 rows per pass; chunks outside [0, stride) -- the 16 bytes left of column 0, or
    // beyond the padded row -- are not loaded (the fix-up below fills them).  Pointers advance by a constant per
    // pass: no address arithmetic inside the loop.
    const uint32_t region_a = smem_addr(region), mid_a = smem_addr(mid);
    for (int c = threadIdx.x & 15; c < nch; c += 16) {            // one trip unless a strong down-scale widens the rows
        const int x = gx0 + 16 * c;
        if (x < 0 || x >= (int)P.src_stride) continue;
        const int r0 = threadIdx.x >> 4;
        const uint8_t* g = src + (size_t)(ry_lo + r0) * P.src_stride + x;
        uint32_t d = region_a + r0 * pitch + c * 16;
        const size_t gstep = (size_t)(kVidThreads / 16) * P.src_stride;
        const uint32_t dstep = (kVidThreads / 16) * pitch;
        for (int r = r0; r < rows; r += kVidThreads / 16, g += gstep, d += dstep) cp_async_16s(d, g);
    }
    // the tile's slices of the tap tables, in flight together with the staging
    if (threadIdx.x < kScaleTW) {
        const uint32_t gx = min(x0 + threadIdx.x, x1);
        const int rel = __ldg(P.xpos + gx) - gx0;                 // >= 0
        const short4 cf = *reinterpret_cast<const short4*>(P.xcoef + (size_t)gx * 4);
        // column x = 4 q + j is kept at [j][q]: the horizontal pass reads one j for 32 consecutive q (no bank conflicts)
        s_x[(threadIdx.x & 3) * (kScaleTW / 4) + (threadIdx.x >> 2)] = make_uint4((uint32_t)(rel & ~3), 0x3210u + 0x1111u * (uint32_t)(rel & 3),
                                      (uint32_t)pack_s16x2(cf.x, cf.y), (uint32_t)pack_s16x2(cf.z, cf.w));
    } else if (threadIdx.x < kScaleTW + kScaleTH) {
        const uint32_t i = threadIdx.x - kScaleTW, gy = min(y0 + i, y1);
        const int p0 = __ldg(P.ypos + gy);
        const short4 cf = *reinterpret_cast<const short4*>(P.ycoef + (size_t)gy * 4);
        s_yr[i] = make_uint4((uint32_t)(min(max(p0, 0), sh - 1) - ry_lo) * kScaleTW, (uint32_t)(min(max(p0 + 1, 0), sh - 1) - ry_lo) * kScaleTW,
                             (uint32_t)(min(max(p0 + 2, 0), sh - 1) - ry_lo) * kScaleTW, (uint32_t)(min(max(p0 + 3, 0), sh - 1) - ry_lo) * kScaleTW);
        s_yc[i] = make_uint2((uint32_t)pack_s16x2(cf.x, cf.y), (uint32_t)pack_s16x2(cf.z, cf.w));
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    // edge tiles: replicate column 0 to the left and column sw-1 to the right (the taps clamp, video scalers do)
    if (gx0 < 0 || gx0 + width > sw) {
        for (int i = threadIdx.x; i < rows * 2; i += kVidThreads) {
            uint8_t* row = region + (i >> 1) * pitch;
            if ((i & 1) == 0) {
                if (gx0 < 0) {
                    const uint8_t v = row[-gx0];
                    for (int c = 0; c < -gx0; c++) row[c] = v;
                }
            } else if (gx0 + width > sw) {
                const uint8_t v = row[sw - 1 - gx0];
                for (int c = sw - gx0; c < width; c++) row[c] = v;
            }
        }
        __syncthreads();
    }
    // horizontal pass: thread = 4 adjacent output columns of one staged row at a time.  The four taps of a pixel are
    // four consecutive staged bytes: two aligned word loads and one byte permute bring them into one register, two
    // dp2a (s16 weights x u8 pixels) reduce them; mid[r][x] packed 4 per store
    {
        const int q = threadIdx.x % (kScaleTW / 4);               // column quad
        uint32_t sel[4], a[4];
        int c01[4], c23[4];
        const int r_first = threadIdx.x / (kScaleTW / 4);
        const uint32_t sx_a = smem_addr(s_x);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint4 e = lds128(sx_a + (j * (kScaleTW / 4) + q) * 16);
            a[j] = region_a + r_first * pitch + e.x;
            sel[j] = e.y; c01[j] = (int)e.z; c23[j] = (int)e.w;
        }
        const uint32_t step = (kVidThreads / (kScaleTW / 4)) * pitch;
        uint32_t out = mid_a + r_first * kScaleTW + 4 * q;
        for (int r = r_first; r < rows; r += kVidThreads / (kScaleTW / 4)) {
            int v[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                v[j] = tap4(__byte_perm(lds32(a[j]), lds32(a[j] + 4), sel[j]), c01[j], c23[j]);
                a[j] += step;
            }
            sts32(out, pack4_sat_u8(v[0], v[1], v[2], v[3]));
            out += (kVidThreads / (kScaleTW / 4)) * kScaleTW;
        }
    }
    __syncthreads();
    // vertical pass: thread = 8 adjacent output columns of one output row at a time, one 64-bit load per tap row.  Two
    // byte permutes per pair of tap rows interleave them so that dp2a.lo / dp2a.hi see (row r, row r+1) of pixel 0 / 1:
    // the 4 x 4 transposition costs four permutes per four pixels, the row table and the weights are read once per eight.
    {
        constexpr int kOct = kScaleTW / 8;                        // column octets per row: 16 threads per output row
        constexpr int kRowsPerPass = kVidThreads / kOct;
        const int q = threadIdx.x % kOct;
        const uint32_t gx = x0 + 8 * q;
        const bool wide_ok = ((P.dst_off | P.dst_stride) & 7u) == 0 && gx + 7 <= x1;
        const uint32_t ly0 = threadIdx.x / kOct;
        uint8_t* o = job.dst + P.dst_off + (size_t)(y0 + ly0) * P.dst_stride + gx;
        const size_t ostep = (size_t)kRowsPerPass * P.dst_stride;
        const uint32_t mq = mid_a + 8 * q;
        uint32_t yr_a = smem_addr(s_yr) + ly0 * 16, yc_a = smem_addr(s_yc) + ly0 * 8;
        for (uint32_t ly = ly0; y0 + ly <= y1; ly += kRowsPerPass, o += ostep, yr_a += kRowsPerPass * 16, yc_a += kRowsPerPass * 8) {
            const uint4 rr = lds128(yr_a);
            const uint2 cc = lds64(yc_a);
            const uint2 a0 = lds64(mq + rr.x), a1 = lds64(mq + rr.y), a2 = lds64(mq + rr.z), a3 = lds64(mq + rr.w);
            const int c01 = (int)cc.x, c23 = (int)cc.y;
            uint32_t out[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint32_t r0 = h ? a0.y : a0.x, r1 = h ? a1.y : a1.x, r2 = h ? a2.y : a2.x, r3 = h ? a3.y : a3.x;
                const uint32_t lo01 = __byte_perm(r0, r1, 0x5140), hi01 = __byte_perm(r0, r1, 0x7362);   // (r0.b0,r1.b0,r0.b1,r1.b1), (.b2,.b3)
                const uint32_t lo23 = __byte_perm(r2, r3, 0x5140), hi23 = __byte_perm(r2, r3, 0x7362);
                const int v0 = dp2a_lo(c23, lo23, dp2a_lo(c01, lo01, 8192)) >> 14, v1 = dp2a_hi(c23, lo23, dp2a_hi(c01, lo01, 8192)) >> 14;
                const int v2 = dp2a_lo(c23, hi23, dp2a_lo(c01, hi01, 8192)) >> 14, v3 = dp2a_hi(c23, hi23, dp2a_hi(c01, hi01, 8192)) >> 14;
                out[h] = pack4_sat_u8(v0, v1, v2, v3);
            }
            if (wide_ok) {
                *reinterpret_cast<uint2*>(o) = make_uint2(out[0], out[1]);
            } else {
                for (uint32_t j = 0; j < 8 && gx + j <= x1; j++) o[j] = (uint8_t)(out[j >> 2] >> (8 * (j & 3)));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Crossfade + colour conversion in one pass (BASELINE config 3: "convert + 2-layer compositor"):
// reads the two yuv420p layers once, blends them with VideoMixer's exact arithmetic (fade4 above) and
// writes RGBA8 of the blend -- 14 515 200 B per 1080p frame instead of 9 331 200 + 11 404 800 for
// crossfade followed by conversion.  One thread = 8 pixels wide x 2 rows (one chroma row).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ldg4(const uint8_t* p) { return __ldg(reinterpret_cast<const uint32_t*>(p)); }
__device__ __forceinline__ uint2 ldg8(const uint8_t* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }

__global__ void __launch_bounds__(kVidThreads) compose_rgba_kernel(const ComposeRgbaJob* __restrict__ jobs, uint32_t width, uint32_t height,
                                                                   uint32_t ystride, uint32_t cstride, uint64_t off_u, uint64_t off_v)
{
    const ComposeRgbaJob job = jobs[blockIdx.z];
    const uint32_t x0 = (blockIdx.x * kVidThreads + threadIdx.x) * 8;
    const uint32_t cy = blockIdx.y;                               // chroma row; luma rows 2cy, 2cy+1
    if (x0 >= width) return;
    const uint32_t f = job.fade, g = 255u - job.fade;
    uint2 ya0 = make_uint2(0u, 0u), ya1 = ya0, yb0 = ya0, yb1 = ya0;
    uint32_t ua = 0x80808080u, va = 0x80808080u, ub = 0x80808080u, vb = 0x80808080u;   // a missing layer is blank
    const uint64_t yo = (uint64_t)(2 * cy) * ystride + x0, co = (uint64_t)cy * cstride + (x0 >> 1);
    const bool row1 = 2 * cy + 1 < height;
    if (job.a) {                                                   // (a layer whose weight is 0 arrives as missing)
        ya0 = ldg8(job.a + yo);
        if (row1) ya1 = ldg8(job.a + yo + ystride);
        ua = ldg4(job.a + off_u + co); va = ldg4(job.a + off_v + co);
    }
    if (job.b) {
        yb0 = ldg8(job.b + yo);
        if (row1) yb1 = ldg8(job.b + yo + ystride);
        ub = ldg4(job.b + off_u + co); vb = ldg4(job.b + off_v + co);
    }
    uint32_t y0w[2], y1w[2], u, v;
    if (g == 0 && job.a) {                                         // (a*255 + b*0) / 255 == a: plain conversion of layer A
        y0w[0] = ya0.x; y0w[1] = ya0.y; y1w[0] = ya1.x; y1w[1] = ya1.y; u = ua; v = va;
    } else if (f == 0 && job.b) {
        y0w[0] = yb0.x; y0w[1] = yb0.y; y1w[0] = yb1.x; y1w[1] = yb1.y; u = ub; v = vb;
    } else {
        y0w[0] = fade4(ya0.x, yb0.x, f, g); y0w[1] = fade4(ya0.y, yb0.y, f, g);
        y1w[0] = fade4(ya1.x, yb1.x, f, g); y1w[1] = fade4(ya1.y, yb1.y, f, g);
        u = fade4(ua, ub, f, g); v = fade4(va, vb, f, g);
    }
    uint32_t px0[8], px1[8];
#pragma unroll
    for (int j = 0; j < 4; j++) {                                  // one chroma sample = a 2x2 block of pixels
        const ChromaTerms ct = chroma_terms(byte_of(u, j), byte_of(v, j));
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int i = 2 * j + h;
            px0[i] = yuv_px_terms(byte_of(y0w[i >> 2], i & 3), ct);
            px1[i] = yuv_px_terms(byte_of(y1w[i >> 2], i & 3), ct);
        }
    }
    uint8_t* o0 = job.rgba + ((uint64_t)(2 * cy) * width + x0) * 4;
    uint8_t* o1 = o0 + (uint64_t)width * 4;
    if (x0 + 8 <= width && (width & 3u) == 0) {
        reinterpret_cast<uint4*>(o0)[0] = make_uint4(px0[0], px0[1], px0[2], px0[3]);
        reinterpret_cast<uint4*>(o0)[1] = make_uint4(px0[4], px0[5], px0[6], px0[7]);
        if (row1) {
            reinterpret_cast<uint4*>(o1)[0] = make_uint4(px1[0], px1[1], px1[2], px1[3]);
            reinterpret_cast<uint4*>(o1)[1] = make_uint4(px1[4], px1[5], px1[6], px1[7]);
        }
    } else {
        for (uint32_t i = 0; i < 8 && x0 + i < width; i++) {
            reinterpret_cast<uint32_t*>(o0)[i] = px0[i];
            if (row1) reinterpret_cast<uint32_t*>(o1)[i] = px1[i];
        }
    }
}

// ------------------------------------------------------------------------------------------
// RGBA8 -> yuv420p, BT.601 limited range (the reference never converts colour, video_mixer.rs:282-283: self-specified,
// oracle = definition).  One thread = 8 pixels x 2 rows = four 2x2 blocks: two 32-byte RGBA row segments in, two 8-byte
// luma segments + 4 U + 4 V bytes out (11.4 MB per 1080p picture: 8.3 read, 3.1 written).  Luma per pixel, chroma from the
// block's rounded mean colour; a block cut by an odd right / bottom edge repeats its last column / row.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t luma_of(uint32_t px)
{
    const int r = px & 255u, g = (px >> 8) & 255u, b = (px >> 16) & 255u;
    return (uint32_t)(((66 * r + 129 * g + 25 * b + 128) >> 8) + 16);
}

__global__ void __launch_bounds__(kVidThreads) rgba_to_yuv_kernel(const RgbaToYuvJob* __restrict__ jobs, uint32_t width, uint32_t height,
                                                                  uint32_t ystride, uint32_t cstride, uint64_t off_u, uint64_t off_v)
{
    const RgbaToYuvJob job = jobs[blockIdx.z];
    const uint32_t x0 = (blockIdx.x * kVidThreads + threadIdx.x) * 8;
    const uint32_t cy = blockIdx.y;
    if (x0 >= width) return;
    const bool row1 = 2 * cy + 1 < height;
    const uint8_t* r0 = job.rgba + ((uint64_t)(2 * cy) * width + x0) * 4;
    const uint8_t* r1 = row1 ? r0 + (uint64_t)width * 4 : r0;          // odd height: the last row again
    uint32_t p0[8], p1[8];
    if (x0 + 8 <= width && (width & 3u) == 0) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(r0)), b = __ldg(reinterpret_cast<const uint4*>(r0) + 1);
        const uint4 c = __ldg(reinterpret_cast<const uint4*>(r1)), d = __ldg(reinterpret_cast<const uint4*>(r1) + 1);
        p0[0] = a.x; p0[1] = a.y; p0[2] = a.z; p0[3] = a.w; p0[4] = b.x; p0[5] = b.y; p0[6] = b.z; p0[7] = b.w;
        p1[0] = c.x; p1[1] = c.y; p1[2] = c.z; p1[3] = c.w; p1[4] = d.x; p1[5] = d.y; p1[6] = d.z; p1[7] = d.w;
    } else {
#pragma unroll
        for (uint32_t i = 0; i < 8; i++) {
            const uint32_t xx = x0 + i < width ? i : width - 1 - x0;    // odd width: the last column again
            p0[i] = __ldg(reinterpret_cast<const uint32_t*>(r0) + xx);
            p1[i] = __ldg(reinterpret_cast<const uint32_t*>(r1) + xx);
        }
    }
    uint32_t y0w[2] = {0u, 0u}, y1w[2] = {0u, 0u}, uw = 0u, vw = 0u;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        y0w[i >> 2] |= luma_of(p0[i]) << (8 * (i & 3));
        y1w[i >> 2] |= luma_of(p1[i]) << (8 * (i & 3));
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t a = p0[2 * j], b = p0[2 * j + 1], c = p1[2 * j], d = p1[2 * j + 1];
        const int r = (int)(((a & 255u) + (b & 255u) + (c & 255u) + (d & 255u) + 2u) >> 2);
        const int g = (int)((((a >> 8) & 255u) + ((b >> 8) & 255u) + ((c >> 8) & 255u) + ((d >> 8) & 255u) + 2u) >> 2);
        const int bl = (int)((((a >> 16) & 255u) + ((b >> 16) & 255u) + ((c >> 16) & 255u) + ((d >> 16) & 255u) + 2u) >> 2);
        uw |= (uint32_t)(((-38 * r - 74 * g + 112 * bl + 128) >> 8) + 128) << (8 * j);
        vw |= (uint32_t)(((112 * r - 94 * g - 18 * bl + 128) >> 8) + 128) << (8 * j);
    }
    uint8_t* yo = job.yuv + (uint64_t)(2 * cy) * ystride + x0;
    uint8_t* uo = job.yuv + off_u + (uint64_t)cy * cstride + (x0 >> 1);
    uint8_t* vo = job.yuv + off_v + (uint64_t)cy * cstride + (x0 >> 1);
    if (x0 + 8 <= width) {
        *reinterpret_cast<uint2*>(yo) = make_uint2(y0w[0], y0w[1]);
        if (row1) *reinterpret_cast<uint2*>(yo + ystride) = make_uint2(y1w[0], y1w[1]);
        *reinterpret_cast<uint32_t*>(uo) = uw;
        *reinterpret_cast<uint32_t*>(vo) = vw;
    } else {
        const uint32_t n = width - x0;
        for (uint32_t i = 0; i < n; i++) {
            yo[i] = (uint8_t)(y0w[i >> 2] >> (8 * (i & 3)));
            if (row1) yo[ystride + i] = (uint8_t)(y1w[i >> 2] >> (8 * (i & 3)));
        }
        for (uint32_t j = 0; j < (n + 1) / 2; j++) { uo[j] = (uint8_t)(uw >> (8 * j)); vo[j] = (uint8_t)(vw >> (8 * j)); }
    }
}

int after_launch(mxl_ctx* ctx, const char* name)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("launch of %s failed: %s", name, cudaGetErrorString(e));
        return MXL_ERR_CUDA;
    }
    ctx->launches++;
    return MXL_OK;
}

int require_device(mxl_ctx* ctx)
{
    if (!ctx || !ctx->has_device()) MXL_FAIL(MXL_ERR_NO_DEVICE, "no CUDA device bound to this context");
    return ctx->activate();
}

struct PlaneGeom { uint32_t rows, padw; };

// video_mixer.rs:176-177 (w >> log2_horz, h >> log2_vert) and the 32-byte `while out < end` of fade_line
PlaneGeom plane_geom(const mxl_frame_layout& lay, int comp)
{
    const uint32_t shift = comp == 0 ? 0 : 1;
    PlaneGeom g;
    g.rows = lay.height >> shift;
    g.padw = (((lay.width >> shift) + 31) / 32) * 32;
    return g;
}

}  // namespace

static bool crossfade_is_flat(const mxl_frame_layout& lay)
{
    for (int c = 0; c < 3; c++) {
        PlaneGeom g = plane_geom(lay, c);
        if (g.padw != lay.stride[c] || g.rows != lay.plane_h[c]) return false;
    }
    return true;
}

int launch_crossfade(mxl_ctx* ctx, const mxl_frame_layout& lay, const FadeJob* jobs_dev, uint32_t n_jobs)
{
    MXL_TRY(require_device(ctx));
    if (n_jobs == 0) return MXL_OK;
    if (crossfade_is_flat(lay)) {
        const uint64_t n16 = lay.size / 16;
        const uint64_t per_block = (uint64_t)kVidThreads * kFadeUnroll;
        dim3 grid((unsigned)((n16 + per_block - 1) / per_block), n_jobs);
        MXL_TIMED(ctx, "crossfade_flat_kernel");
        crossfade_flat_kernel<<<grid, kVidThreads, 0, ctx->stream>>>(jobs_dev, n16, lay.offset[1] / 16);
        return after_launch(ctx, "crossfade_flat_kernel");
    }
    for (int c = 0; c < 3; c++) {
        PlaneGeom g = plane_geom(lay, c);
        if (g.rows == 0 || g.padw == 0) continue;
        const uint32_t vpr = g.padw / 16;
        dim3 grid((vpr + kVidThreads - 1) / kVidThreads, g.rows, n_jobs);
        MXL_TIMED(ctx, "crossfade_plane_kernel");
        crossfade_plane_kernel<<<grid, kVidThreads, 0, ctx->stream>>>(jobs_dev, lay.offset[c], lay.stride[c], vpr,
                                                                     c == 0 ? 0u : 0x80808080u);
        MXL_TRY(after_launch(ctx, "crossfade_plane_kernel"));
    }
    return MXL_OK;
}

int launch_crossfade_inline(mxl_ctx* ctx, const mxl_frame_layout& lay, const FadeJob* jobs_host, uint32_t n_jobs)
{
    MXL_TRY(require_device(ctx));
    if (n_jobs == 0) return MXL_OK;
    if (n_jobs > (uint32_t)kFadeInlineJobs) MXL_FAIL(MXL_ERR_INVALID, "launch_crossfade_inline: %u jobs, at most %d", n_jobs, kFadeInlineJobs);
    FadeJobsInline tab;
    for (uint32_t i = 0; i < n_jobs; i++) tab.job[i] = jobs_host[i];
    if (crossfade_is_flat(lay)) {
        const uint64_t n16 = lay.size / 16;
        const uint64_t per_block = (uint64_t)kVidThreads * kFadeUnroll;
        dim3 grid((unsigned)((n16 + per_block - 1) / per_block), n_jobs);
        MXL_TIMED(ctx, "crossfade_flat_kernel");
        crossfade_flat_inline_kernel<<<grid, kVidThreads, 0, ctx->stream>>>(tab, n16, lay.offset[1] / 16);
        return after_launch(ctx, "crossfade_flat_inline_kernel");
    }
    for (int c = 0; c < 3; c++) {
        PlaneGeom g = plane_geom(lay, c);
        if (g.rows == 0 || g.padw == 0) continue;
        const uint32_t vpr = g.padw / 16;
        dim3 grid((vpr + kVidThreads - 1) / kVidThreads, g.rows, n_jobs);
        MXL_TIMED(ctx, "crossfade_plane_kernel");
        crossfade_plane_inline_kernel<<<grid, kVidThreads, 0, ctx->stream>>>(tab, lay.offset[c], lay.stride[c], vpr,
                                                                            c == 0 ? 0u : 0x80808080u);
        MXL_TRY(after_launch(ctx, "crossfade_plane_inline_kernel"));
    }
    return MXL_OK;
}

int launch_blank(mxl_ctx* ctx, const mxl_frame_layout& lay, uint8_t* frame)
{
    MXL_TRY(require_device(ctx));
    const uint64_t n16 = lay.size / 16;
    unsigned blocks = (unsigned)((n16 + kVidThreads - 1) / kVidThreads);
    const unsigned cap = ctx->sm_count > 0 ? ctx->sm_count * 8 : 1184;
    if (blocks > cap) blocks = cap;
    if (blocks == 0) return MXL_OK;
    MXL_TIMED(ctx, "blank_kernel");
    blank_kernel<<<blocks, kVidThreads, 0, ctx->stream>>>(reinterpret_cast<uint4*>(frame), n16, lay.offset[1] / 16);
    return after_launch(ctx, "blank_kernel");
}

template <int TH>
static int launch_scale_th(mxl_ctx* ctx, const ScaleLaunch& L, uint32_t n_jobs, size_t smem)
{
    uint32_t& configured = ctx->scale_smem[TH == 128 ? 4 : (TH == 64 ? 3 : (TH == 32 ? 0 : (TH == 8 ? 1 : 2)))];
    if (smem > 48 * 1024 && smem > configured) {
        MXL_CUDA(cudaFuncSetAttribute(scale_tiled_kernel<TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = (uint32_t)smem;
    }
    dim3 grid(L.total_tiles, n_jobs);
    MXL_TIMED(ctx, "scale_tiled_kernel");
    scale_tiled_kernel<TH><<<grid, kVidThreads, smem, ctx->stream>>>(L);
    return after_launch(ctx, "scale_tiled_kernel");
}

size_t scale_smem_bytes(uint32_t region_rows, uint32_t region_pitch) { return (size_t)region_rows * region_pitch + (size_t)region_rows * kScaleTW; }
uint32_t scale_tile_width() { return kScaleTW; }

int launch_scale_tiled(mxl_ctx* ctx, const ScaleLaunch& L, uint32_t n_jobs)
{
    MXL_TRY(require_device(ctx));
    if (n_jobs == 0 || L.total_tiles == 0) return MXL_OK;
    const size_t smem = scale_smem_bytes(L.region_rows, L.region_pitch);
    if (smem > kScaleMaxSmem) MXL_FAIL(MXL_ERR_INVALID, "scale_tiled_kernel: source rectangle of a tile needs %zu bytes of shared memory", smem);
    switch (L.tile_h) {
    case 128: return launch_scale_th<128>(ctx, L, n_jobs, smem);
    case 64: return launch_scale_th<64>(ctx, L, n_jobs, smem);
    case 32: return launch_scale_th<32>(ctx, L, n_jobs, smem);
    case 8: return launch_scale_th<8>(ctx, L, n_jobs, smem);
    case 2: return launch_scale_th<2>(ctx, L, n_jobs, smem);
    default: MXL_FAIL(MXL_ERR_INVALID, "scale_tiled_kernel: tile height %u", L.tile_h);
    }
}

int launch_compose_rgba(mxl_ctx* ctx, const mxl_frame_layout& lay, const ComposeRgbaJob* jobs_dev, uint32_t n_jobs)
{
    MXL_TRY(require_device(ctx));
    if (n_jobs == 0 || lay.width == 0 || lay.height == 0) return MXL_OK;
    if ((lay.stride[0] & 7) || (lay.stride[1] & 3) || (lay.offset[1] & 3) || (lay.offset[2] & 3))
        MXL_FAIL(MXL_ERR_INVALID, "compose_rgba: plane strides/offsets must be 8/4-byte aligned");
    dim3 grid(((lay.width + 7) / 8 + kVidThreads - 1) / kVidThreads, (lay.height + 1) / 2, n_jobs);
    MXL_TIMED(ctx, "compose_rgba_kernel");
    compose_rgba_kernel<<<grid, kVidThreads, 0, ctx->stream>>>(jobs_dev, lay.width, lay.height, lay.stride[0], lay.stride[1],
                                                               lay.offset[1], lay.offset[2]);
    return after_launch(ctx, "compose_rgba_kernel");
}

int launch_rgba_to_yuv(mxl_ctx* ctx, const mxl_frame_layout& lay, const RgbaToYuvJob* jobs_dev, uint32_t n_jobs)
{
    if (!ctx || !ctx->has_device()) MXL_FAIL(MXL_ERR_NO_DEVICE, "no CUDA device bound to this context");
    MXL_TRY(ctx->activate());
    if (n_jobs == 0) return MXL_OK;
    if ((lay.stride[0] & 7) || (lay.stride[1] & 3) || (lay.offset[1] & 3) || (lay.offset[2] & 3))
        MXL_FAIL(MXL_ERR_INVALID, "rgba_to_yuv: plane strides/offsets must be 8/4-byte aligned");
    dim3 grid(((lay.width + 7) / 8 + kVidThreads - 1) / kVidThreads, (lay.height + 1) / 2, n_jobs);
    MXL_TIMED(ctx, "rgba_to_yuv_kernel");
    rgba_to_yuv_kernel<<<grid, kVidThreads, 0, ctx->stream>>>(jobs_dev, lay.width, lay.height, lay.stride[0], lay.stride[1],
                                                              lay.offset[1], lay.offset[2]);
    return after_launch(ctx, "rgba_to_yuv_kernel");
}

}  // namespace k
}  // namespace mxl
