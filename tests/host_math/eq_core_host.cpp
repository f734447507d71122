// Host build of mixlab_b200/csrc/eq_core.cuh + eq_plan.h for CPU-side verification (tests only):
// the sequential reference form, the skewed chunk runner, and a CPU model of the whole
// time-parallel scheme of eq_stream.cu (dot-product zero-state pass, windowed scan, skewed re-run).
#include "../../mixlab_b200/csrc/eq_core.cuh"
#include "../../mixlab_b200/csrc/eq_plan.h"
#include <stddef.h>
#include <vector>

namespace {
struct HostIo {
    float* p;
    mxl::EqF4 load(int v) const { return mxl::EqF4{p[4 * v], p[4 * v + 1], p[4 * v + 2], p[4 * v + 3]}; }
    void store(int v, mxl::EqF4 y) { p[4 * v] = y.x; p[4 * v + 1] = y.y; p[4 * v + 2] = y.z; p[4 * v + 3] = y.w; }
};
struct HostTab {
    const mxl::EqStreamPlan* pl;
    double v(int j, int e) const { return pl->V[j][e]; }
};
void tri_apply(const double* A, const double* x, double* y)
{
    for (int r = 0; r < 4; r++) {
        double acc = 0.0;
        for (int c = 0; c <= r; c++) acc = fma(A[r * (r + 1) / 2 + c], x[c], acc);
        y[r] = acc;
    }
}
mxl::EqGains gains(unsigned sr, const double* g3)
{
    mxl::EqGains g;
    mxl::eq_coefficients(sr, &g.cl, &g.ch);
    g.g_lo = g3[0]; g.g_mid = g3[1]; g.g_hi = g3[2];
    return g;
}
template <int LC>
void run_skewed(mxl::EqPoles& p, double* hist, float* buf, const mxl::EqGains& g)
{
    HostIo io{buf};
    mxl::eq_run_chunk_skewed<LC>(p, hist, io, g);
}
void run_chunk(int lc, mxl::EqPoles& p, double* hist, float* buf, const mxl::EqGains& g)
{
    if (lc == 8) run_skewed<8>(p, hist, buf, g);
    else if (lc == 16) run_skewed<16>(p, hist, buf, g);
    else if (lc == 32) run_skewed<32>(p, hist, buf, g);
    else run_skewed<64>(p, hist, buf, g);
}
template <int LC>
void zero_dot(const float* buf, const mxl::EqStreamPlan& pl, double* acc)
{
    HostIo io{const_cast<float*>(buf)};
    HostTab tab{&pl};
    mxl::eq_zero_state_dot<LC>(io, tab, acc);
}
}  // namespace

extern "C" {

// state = lo poles[4], hi poles[4], history[3]; updated in place
void mxl_host_eq_seq(const float* in, float* out, size_t n, double* state, unsigned sr, const double* g3)
{
    const mxl::EqGains g = gains(sr, g3);
    mxl::EqPoles p{state[0], state[1], state[2], state[3], state[4], state[5], state[6], state[7]};
    double hist[3] = {state[8], state[9], state[10]};
    for (size_t i = 0; i < n; i++) out[i] = mxl::eq_step_seq(p, hist, in[i], g);
    const double s[11] = {p.l0, p.l1, p.l2, p.l3, p.h0, p.h1, p.h2, p.h3, hist[0], hist[1], hist[2]};
    for (int i = 0; i < 11; i++) state[i] = s[i];
}

// chunks of lc samples run by the skewed runner from the exactly carried state (tail sequential)
void mxl_host_eq_skewed(const float* in, float* out, size_t n, double* state, unsigned sr, const double* g3, int lc)
{
    const mxl::EqGains g = gains(sr, g3);
    mxl::EqPoles p{state[0], state[1], state[2], state[3], state[4], state[5], state[6], state[7]};
    double hist[3] = {state[8], state[9], state[10]};
    size_t i = 0;
    for (; i + lc <= n; i += lc) {
        for (int j = 0; j < lc; j++) out[i + j] = in[i + j];
        run_chunk(lc, p, hist, out + i, g);
    }
    for (; i < n; i++) out[i] = mxl::eq_step_seq(p, hist, in[i], g);
    const double s[11] = {p.l0, p.l1, p.l2, p.l3, p.h0, p.h1, p.h2, p.h3, hist[0], hist[1], hist[2]};
    for (int k = 0; k < 11; k++) state[k] = s[k];
}

int mxl_host_eq_plan(unsigned sr, unsigned lc, unsigned max_halo, mxl::EqStreamPlan* out)
{
    *out = mxl::eq_stream_plan(sr, lc, max_halo);
    return out->ok ? 1 : 0;
}
size_t mxl_host_eq_plan_size() { return sizeof(mxl::EqStreamPlan); }
unsigned mxl_host_eq_plan_field(const mxl::EqStreamPlan* p, int which)
{
    return which == 0 ? p->halo : (which == 1 ? p->lev_lo : p->lev_hi);
}

// CPU model of eq_stream_kernel, CTA by CTA and warp by warp exactly as the kernel forms its carries:
// 256 chunks per CTA of which the first `halo` only feed the scan, shuffle scan inside each 32-chunk
// warp, warp aggregates + per-lane powers across warps.
int mxl_host_eq_parallel(const float* in, float* out, size_t n, double* state, unsigned sr, const double* g3, int lc)
{
    const mxl::EqStreamPlan pl = mxl::eq_stream_plan(sr, (unsigned)lc, 128);
    if (!pl.ok) return 0;
    const mxl::EqGains g = gains(sr, g3);
    const long nc = (long)((n + lc - 1) / lc);
    const int T = 256, halo = (int)pl.halo, U = T - halo;
    double fin_state[11];
    for (int k = 0; k < 11; k++) fin_state[k] = state[k];
    std::vector<float> pad(lc);
    for (long blk = 0; blk * U < nc; blk++) {
        const long c0 = blk * U - halo;
        double v[256][8];
        for (int tid = 0; tid < T; tid++) {
            const long c = c0 + tid;
            for (int e = 0; e < 8; e++) v[tid][e] = 0.0;
            if (c < 0 || c >= nc) continue;
            for (int j = 0; j < lc; j++) pad[j] = (size_t)(c * lc + j) < n ? in[c * lc + j] : 0.f;
            double acc[8];
            for (int e = 0; e < 8; e++) acc[e] = pl.K[e];
            if (lc == 8) zero_dot<8>(pad.data(), pl, acc);
            else if (lc == 16) zero_dot<16>(pad.data(), pl, acc);
            else if (lc == 32) zero_dot<32>(pad.data(), pl, acc);
            else zero_dot<64>(pad.data(), pl, acc);
            if (c == 0) {
                double yl[4], yh[4];
                tri_apply(pl.pow_lo[0], state, yl);
                tri_apply(pl.pow_hi[0], state + 4, yh);
                for (int e = 0; e < 4; e++) { acc[e] += yl[e]; acc[4 + e] += yh[e]; }
            }
            for (int e = 0; e < 8; e++) v[tid][e] = acc[e];
        }
        // shuffle scan inside each warp
        for (int d = 0; d < 5; d++) {
            const bool lo_live = d < (int)pl.lev_lo, hi_live = d < (int)pl.lev_hi;
            if (!lo_live && !hi_live) break;
            double nv[256][8];
            memcpy(nv, v, sizeof v);
            for (int tid = 0; tid < T; tid++) {
                const int lane = tid & 31;
                if (lane < (1 << d)) continue;
                double y[4];
                if (lo_live) { tri_apply(pl.pow_lo[d], v[tid - (1 << d)], y); for (int e = 0; e < 4; e++) nv[tid][e] = v[tid][e] + y[e]; }
                if (hi_live) { tri_apply(pl.pow_hi[d], v[tid - (1 << d)] + 4, y); for (int e = 0; e < 4; e++) nv[tid][4 + e] = v[tid][4 + e] + y[e]; }
            }
            memcpy(v, nv, sizeof v);
        }
        double agg[8][8];
        for (int w = 0; w < 8; w++) for (int e = 0; e < 8; e++) agg[w][e] = v[w * 32 + 31][e];
        for (int tid = 32; tid < T; tid++) {
            const int lane = tid & 31, warp = tid >> 5;
            double P[8];
            for (int e = 0; e < 8; e++) P[e] = agg[warp - 1][e];
            for (int k = 1; k <= 2; k++) {
                if (warp - 1 - k < 0) break;
                const bool lo_live = k < (int)pl.back_lo, hi_live = k < (int)pl.back_hi;
                if (!lo_live && !hi_live) break;
                double y[4];
                if (lo_live) { tri_apply(pl.pow_lo[4 + k], agg[warp - 1 - k], y); for (int e = 0; e < 4; e++) P[e] += y[e]; }
                if (hi_live) { tri_apply(pl.pow_hi[4 + k], agg[warp - 1 - k] + 4, y); for (int e = 0; e < 4; e++) P[4 + e] += y[e]; }
            }
            double Al[10], Ah[10], y[4];
            for (int q = 0; q < 10; q++) { Al[q] = pl.lane_pow[0][q][lane]; Ah[q] = pl.lane_pow[1][q][lane]; }
            tri_apply(Al, P, y);
            for (int e = 0; e < 4; e++) v[tid][e] += y[e];
            tri_apply(Ah, P + 4, y);
            for (int e = 0; e < 4; e++) v[tid][4 + e] += y[e];
        }
        // exact re-run of the owned chunks
        for (int tid = halo; tid < T; tid++) {
            const long c = c0 + tid;
            if (c < 0 || c >= nc) continue;
            const double* S = c == 0 ? state : v[tid - 1];
            mxl::EqPoles p{S[0], S[1], S[2], S[3], S[4], S[5], S[6], S[7]};
            double hist[3];
            for (int j = 0; j < 3; j++) {
                const long idx = c * lc - 3 + j;
                hist[j] = idx >= 0 ? (double)in[idx] : state[8 + 3 + idx];
            }
            const size_t s0 = (size_t)c * lc, cnt = s0 + lc <= n ? (size_t)lc : n - s0;
            if (cnt == (size_t)lc) {
                for (int j = 0; j < lc; j++) out[s0 + j] = in[s0 + j];
                run_chunk(lc, p, hist, out + s0, g);
            } else {
                for (size_t j = 0; j < cnt; j++) out[s0 + j] = mxl::eq_step_seq(p, hist, in[s0 + j], g);
            }
            if (c + 1 == nc) {
                const double s[11] = {p.l0, p.l1, p.l2, p.l3, p.h0, p.h1, p.h2, p.h3, hist[0], hist[1], hist[2]};
                for (int k = 0; k < 11; k++) fin_state[k] = s[k];
            }
        }
    }
    for (int k = 0; k < 11; k++) state[k] = fin_state[k];
    return 1;
}
}
