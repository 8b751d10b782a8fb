// vp.cu -- a8-a12: tracks -> flow lines -> cross points -> vanishing point, one frame per launch, one CTA per clip.
//
//   a8  checkInside + `st[~filter] = 0` incl. the dtype flip          VP.py:326-349, 552-553   
//   a9  FlowLine geometry (float64; length rounded to 2 dp; angle)     VP.py:109-125, 276-283
//   a10 line acceptance with the per-group length EMA                  VP.py:565-574
//   a11 all C(L,2) pairs in itertools.combinations order, static filters, float32 intersection
//                                                                      VP.py:296-324, 581-595   
//   a12 VP gate / init (aliasing quirk) / +-1 sigma inlier update / hide, CSV rows
//                                                                      VP.py:157-194, 596-653   
//   plus the refresh decision of VP.py:700 and the REPlace of VP.py:713-715 (k_apply_refresh), and the vanishing-line
//   pair of class VL (VP.py:196-250, updated at VP.py:654; drawing only in the scripts): two least-squares fits over the
//   last VP_REF history rows, summed by the whole CTA.
//
// Pair geometry is evaluated by the whole CTA (ballot + prefix compaction keeps combinations order); the
// vanishing-point recurrence itself is strictly sequential (the gate uses the current VP) and runs on
// one thread in float64 with explicitly rounded operations, in the summation order numpy uses
// (np.mean / np.std over axis 0 are sequential sums; np.dot of a 2-vector is fma(v1,v1,v0*v0)).
// Oracle: oracle/vp_oracle.py (validated byte-for-byte against the unmodified scripts' CSVs).
#include "common.cuh"
#include <math.h>

namespace {

constexpr int VP_NT = 128;

struct LineRec {
    float sx, sy, ex, ey;      // start (old point), stop (new point), float32 like the LK outputs
    double angle;
};

struct CpRec { float x, y; };

__device__ __forceinline__ double np_round2(double v) { return __ddiv_rn(rint(__dmul_rn(v, 100.0)), 100.0); }

// one accepted cross point: gate, append, update or init (sequential semantics, thread 0 only)
__device__ void consume_cp(lkf_vp_state &S, const lkf_vp_params &P, float x, float y, double *rows, int row_cap)
{
    if (S.vp_init) {
        double ax = fabs(__dsub_rn(S.vp_x, (double)x)), ay = fabs(__dsub_rn(S.vp_y, (double)y));
        if (!(ax < P.cp_gate_x && ay < P.cp_gate_y)) return;
    }
    const int R = P.vp_ref_num;
    int slot = S.ring_head;
    S.ring_x[slot] = (double)x; S.ring_y[slot] = (double)y; S.ring_alias[slot] = 0;
    S.ring_head = (slot + 1) % LKF_VP_RING;
    if (S.ring_count < 0x7fffffff) S.ring_count++;
    S.n_cps++;
    const int n = S.ring_count < R ? S.ring_count : R;
    if (S.vp_init) {
        double dx[LKF_VP_RING], dy[LKF_VP_RING];
        double mx = 0.0, my = 0.0;
        for (int k = 0; k < n; k++) {                       // chronological: oldest of the last n first
            int s = (S.ring_head - n + k + 2 * LKF_VP_RING) % LKF_VP_RING;
            double cx = S.ring_alias[s] ? S.vp_x : S.ring_x[s], cy = S.ring_alias[s] ? S.vp_y : S.ring_y[s];
            dx[k] = __dsub_rn(cx, S.vp_x); dy[k] = __dsub_rn(cy, S.vp_y);
            mx = __dadd_rn(mx, dx[k]); my = __dadd_rn(my, dy[k]);
        }
        mx = __ddiv_rn(mx, (double)n); my = __ddiv_rn(my, (double)n);
        double vx = 0.0, vy = 0.0;
        for (int k = 0; k < n; k++) {
            double ex = __dsub_rn(dx[k], mx), ey = __dsub_rn(dy[k], my);
            vx = __dadd_rn(vx, __dmul_rn(ex, ex)); vy = __dadd_rn(vy, __dmul_rn(ey, ey));
        }
        const double sx = __dmul_rn(__dsqrt_rn(__ddiv_rn(vx, (double)n)), P.max_cp_std);
        const double sy = __dmul_rn(__dsqrt_rn(__ddiv_rn(vy, (double)n)), P.max_cp_std);
        const double hix = __dadd_rn(mx, sx), hiy = __dadd_rn(my, sy), lox = __dsub_rn(mx, sx), loy = __dsub_rn(my, sy);
        double ax = 0.0, ay = 0.0;
        int c = 0;
        for (int k = 0; k < n; k++)
            if (dx[k] <= hix && dy[k] <= hiy && dx[k] >= lox && dy[k] >= loy) {
                ax = __dadd_rn(ax, dx[k]); ay = __dadd_rn(ay, dy[k]); c++;
            }
        if (c) {
            ax = __ddiv_rn(ax, (double)c); ay = __ddiv_rn(ay, (double)c);
            S.vp_x = __dadd_rn(S.vp_x, __dmul_rn(ax, P.vp_update_rate));
            S.vp_y = __dadd_rn(S.vp_y, __dmul_rn(ay, P.vp_update_rate));
            S.vp_moved = 1;                                    // Point.update sets _moved (VP.py:180)
            if (S.n_rows < row_cap) { rows[2 * S.n_rows] = S.vp_x; rows[2 * S.n_rows + 1] = S.vp_y; S.n_rows++; }
            else S.row_overflow = 1;
            S.vp_ult = 0;
        }
    } else if (S.ring_count >= R) {
        double ax = 0.0, ay = 0.0;
        const int m = S.ring_count < LKF_VP_RING ? S.ring_count : LKF_VP_RING;      // == R (init fires at R)
        for (int k = 0; k < m; k++) {
            int s = (S.ring_head - m + k + 2 * LKF_VP_RING) % LKF_VP_RING;
            ax = __dadd_rn(ax, S.ring_x[s]); ay = __dadd_rn(ay, S.ring_y[s]);
        }
        S.vp_x = __ddiv_rn(ax, (double)R); S.vp_y = __ddiv_rn(ay, (double)R);
        S.ring_alias[slot] = 1;                               // the last CP object BECOMES the VP (B.2)
        S.vp_init = 1;
        S.vp_moved = 0;                                       // ... an object that was never update()d
        S.vp_ult = 0;
    }
}

__global__ void __launch_bounds__(VP_NT)
k_vp_frame(lkf_vp_params P, lkf_vp_state *__restrict__ states, float *__restrict__ p0_all, const float *__restrict__ p1_all,
           const uint8_t *__restrict__ st_all, int32_t *__restrict__ counts, int maxp, const uint8_t *__restrict__ mask,
           long long mask_rp, int mH, int mW, double *__restrict__ rows_all, int row_cap, int32_t *__restrict__ refresh_out,
           const int32_t *__restrict__ clip_enable)
{
    extern __shared__ __align__(16) unsigned char vp_smem[];
    const int b = blockIdx.x, tid = threadIdx.x;
    if (clip_enable && !clip_enable[b]) return;
    LineRec *lines = reinterpret_cast<LineRec *>(vp_smem);                 // [2*maxp]
    float *tp0 = reinterpret_cast<float *>(lines + 2 * maxp);             // [2][maxp][2] compacted old points
    float *tp1 = tp0 + 4 * maxp;                                          // [2][maxp][2] compacted new points
    double *lens = reinterpret_cast<double *>(tp1 + 4 * maxp);            // [maxp] rounded line lengths of one group
    int *flag = reinterpret_cast<int *>(lens + maxp);                     // [maxp] scratch
    CpRec *cps = reinterpret_cast<CpRec *>(flag + maxp + (maxp & 1));     // [VP_NT] survivors of one pair chunk
    __shared__ int s_k[2], s_any_none, s_any_one, s_any_zero, s_nl, s_ncp, s_scan[VP_NT / 32], s_do_vl;
    __shared__ double s_red[5][VP_NT / 32], s_mean[2];
    __shared__ lkf_vp_state S;
    if (tid == 0) S = states[b];
    double *rows = rows_all + (size_t)b * row_cap * 2;
    const int n_in[2] = {counts[2 * b], counts[2 * b + 1]};
    if (tid == 0) { s_k[0] = n_in[0]; s_k[1] = n_in[1]; s_nl = 0; }
    __syncthreads();
    const bool any_pts = n_in[0] != 0 || n_in[1] != 0;

    if (any_pts) {
        for (int g = 0; g < 2; g++) {
            const int n = n_in[g];
            if (n == 0) continue;                                              // VP.py:539 `continue`
            const float *p0 = p0_all + ((size_t)(2 * b + g) * maxp) * 2;
            const float *p1 = p1_all + ((size_t)(2 * b + g) * maxp) * 2;
            const uint8_t *st = st_all + (size_t)(2 * b + g) * maxp;
            if (tid == 0) { s_any_none = 0; s_any_one = 0; s_any_zero = 0; }
            __syncthreads();
            // a8: per-track flag  -1 = the int 0 the script appends, 0/1 = mask lookup
            for (int i = tid; i < n; i += VP_NT) {
                float x = p1[2 * i], y = p1[2 * i + 1];
                int fx = (int)floorf(x), fy = (int)floorf(y), f;
                if (st[i] == 0 || fy > mH || fx > mW || !isfinite(x) || !isfinite(y)) f = -1;
                else {
                    int yy = fy < 0 ? fy + mH : fy, xx = fx < 0 ? fx + mW : fx;      // python negative index wrap
                    f = (yy >= 0 && yy < mH && xx >= 0 && xx < mW) ? (mask[(long long)yy * mask_rp + xx] > 0) : 0;
                }
                flag[i] = f;
                if (f < 0) { atomicOr(&s_any_none, 1); atomicOr(&s_any_zero, 1); }
                else if (f == 1) atomicOr(&s_any_one, 1);
                else atomicOr(&s_any_zero, 1);
            }
            __syncthreads();
            // final status, then ordered compaction (thread 0: n is small; keeps the exact order)
            if (tid == 0) {
                int k = 0;
                for (int i = 0; i < n; i++) {
                    int s = st[i];
                    if (!s_any_none) { if (flag[i] == 0) s = 0; }
                    else {
                        if (s_any_zero && i == n - 1) s = 0;
                        if (s_any_one && i == n - 2) s = 0;
                    }
                    if (s == 1) {
                        tp0[(g * maxp + k) * 2] = p0[2 * i]; tp0[(g * maxp + k) * 2 + 1] = p0[2 * i + 1];
                        tp1[(g * maxp + k) * 2] = p1[2 * i]; tp1[(g * maxp + k) * 2 + 1] = p1[2 * i + 1];
                        k++;
                    }
                }
                s_k[g] = k;
            }
            __syncthreads();
            // a9 geometry in parallel -> candidate records + rounded lengths; a10 (EMA) is sequential below
            const int k = s_k[g];
            LineRec *cand = lines + s_nl;                                      // candidates sit after the accepted lines
            for (int i = tid; i < k; i += VP_NT) {
                float c = tp0[(g * maxp + i) * 2], d = tp0[(g * maxp + i) * 2 + 1];
                float a = tp1[(g * maxp + i) * 2], e = tp1[(g * maxp + i) * 2 + 1];
                LineRec r;
                r.sx = c; r.sy = d; r.ex = a; r.ey = e;
                double v0 = __dsub_rn((double)a, (double)c);                       // stop*[1,-1] - start*[1,-1]
                double v1 = __dsub_rn(__dmul_rn((double)e, -1.0), __dmul_rn((double)d, -1.0));
                double nrm = __dsqrt_rn(__fma_rn(v1, v1, __dmul_rn(v0, v0)));      // np.dot of a 2-vector (OpenBLAS: fma)
                double u0 = __ddiv_rn(v0, nrm);
                double cc = fmin(fmax(u0, -1.0), 1.0);
                if (u0 != u0) cc = u0;                                              // np.clip keeps NaN
                double ang = __dmul_rn(__ddiv_rn(acos(cc), 3.141592653589793), 180.0);
                if (v1 < 0) ang = __dsub_rn(360.0, ang);
                r.angle = ang;
                cand[i] = r;
                lens[i] = (a == c && e == d) ? -1.0 : np_round2(nrm);              // -1: "not moved", skipped
            }
            __syncthreads();
            if (tid == 0) {
                int nl = s_nl;
                for (int i = 0; i < k; i++) {
                    LineRec r = cand[i];
                    const double L = lens[i];
                    if (L < 0.0) continue;                                          // a == c and b == d
                    if (r.angle > 180.0 && L > P.min_fl_len) {
                        S.avg_len[g] = __ddiv_rn(__dadd_rn(S.avg_len[g], __dmul_rn(L, P.fl_update_rate)),
                                                 __dadd_rn(1.0, P.fl_update_rate));
                        if (L > S.avg_len[g]) lines[nl++] = r;                      // nl <= s_nl + i: never overtakes cand[i]
                    }
                }
                S.n_lines += nl - s_nl;
                s_nl = nl;
            }
            __syncthreads();
        }
        // a11 + a12: pairs in combinations order, chunks of VP_NT pairs
        const int L = s_nl;
        const long long npairs = (long long)L * (L - 1) / 2;
        for (long long base = 0; base < npairs; base += VP_NT) {
            long long kp = base + tid;
            bool ok = false;
            float cx = 0.f, cy = 0.f;
            if (kp < npairs) {
                // pair rank kp -> (i, j), i < j, lexicographic
                double Ld = (double)L;
                long long i = (long long)floor(((2.0 * Ld - 1.0) - sqrt((2.0 * Ld - 1.0) * (2.0 * Ld - 1.0) - 8.0 * (double)kp)) * 0.5);
                if (i < 0) i = 0;
                while (i > 0 && i * (2 * L - i - 1) / 2 > kp) i--;
                while ((i + 1) * (2 * L - (i + 1) - 1) / 2 <= kp) i++;
                long long j = kp - i * (2 * L - i - 1) / 2 + i + 1;
                const LineRec la = lines[i], lb = lines[j];
                double ad = fabs(__dsub_rn(la.angle, lb.angle));
                ok = !(ad < P.min_ang_dif || ad > __dsub_rn(360.0, P.min_ang_dif));
                if (ok && P.start_x_sep >= 0.f) ok = !(fabsf(__fsub_rn(la.sx, lb.sx)) < P.start_x_sep);
                if (ok) {
                    // cross_point(line1 = lb, line2 = la), float32 scalar arithmetic (numpy >= 2)
                    float x1 = lb.sx, y1 = lb.sy, x2 = lb.ex, y2 = lb.ey, x3 = la.sx, y3 = la.sy, x4 = la.ex, y4 = la.ey;
                    float k1 = __fdiv_rn(__fsub_rn(y2, y1), __fsub_rn(x2, x1));
                    float b1 = __fsub_rn(y1, __fmul_rn(x1, k1));
                    float x;
                    if (__fsub_rn(x4, x3) == 0.f) x = x3;
                    else {
                        float k2 = __fdiv_rn(__fsub_rn(y4, y3), __fsub_rn(x4, x3));
                        float b2 = __fsub_rn(y3, __fmul_rn(x3, k2));
                        float dk = __fsub_rn(k1, k2);
                        if (dk == 0.f) ok = false;
                        x = __fdiv_rn(__fsub_rn(b2, b1), dk);
                    }
                    float y = __fadd_rn(__fmul_rn(k1, x), b1);
                    if (ok && (y > la.sy || y > lb.sy)) ok = false;
                    cx = x; cy = y;
                }
            }
            // ordered compaction of this chunk's survivors
            unsigned bal = __ballot_sync(0xffffffffu, ok);
            int lane = tid & 31, wid = tid >> 5;
            if (lane == 0) s_scan[wid] = __popc(bal);
            __syncthreads();
            int off = 0, tot = 0;
            for (int w = 0; w < VP_NT / 32; w++) { int c = s_scan[w]; if (w < wid) off += c; tot += c; }
            if (ok) { CpRec r; r.x = cx; r.y = cy; cps[off + __popc(bal & ((1u << lane) - 1))] = r; }
            if (tid == 0) s_ncp = tot;
            __syncthreads();
            if (tid == 0)
                for (int q = 0; q < s_ncp; q++) consume_cp(S, P, cps[q].x, cps[q].y, rows, row_cap);
            __syncthreads();
        }
    }
    // Step 8 (hide / per-frame row), track hand-over, refresh decision, counters
    const int k0 = any_pts && n_in[0] ? s_k[0] : n_in[0], k1 = any_pts && n_in[1] ? s_k[1] : n_in[1];
    if (any_pts)
        for (int g = 0; g < 2; g++) {
            if (n_in[g] == 0) continue;
            float *p0 = p0_all + ((size_t)(2 * b + g) * maxp) * 2;
            const int k = s_k[g];
            for (int i = tid; i < 2 * k; i += VP_NT) p0[i] = tp1[g * maxp * 2 + i];       // p0s[n] = good_new
        }
    if (tid == 0) {
        s_do_vl = 0;
        if (S.vp_init) {
            if (S.vp_ult > P.hide_vp_thold) {
                S.vp_init = 0; S.ring_count = 0; S.ring_head = 0; S.vp_moved = 0;
                for (int q = 0; q < LKF_VP_RING; q++) S.ring_alias[q] = 0;
                if (P.reset_avg_on_hide) { S.avg_len[0] = P.min_fl_len; S.avg_len[1] = P.min_fl_len; }
            } else {
                if (S.n_rows < row_cap) { rows[2 * S.n_rows] = S.vp_x; rows[2 * S.n_rows + 1] = S.vp_y; S.n_rows++; }
                else S.row_overflow = 1;
                s_do_vl = S.vp_moved && P.vl_ref >= 0 && !S.row_overflow;     // vl.update(all_vp[-VP_REF:], vp), VP.py:654
            }
        }
    }
    __syncthreads();
    if (s_do_vl) {
        // VL.update (VP.py:224-245): scipy.stats.linregress slope = (dx.dy / n) / (dx.dx / n) on the de-meaned samples,
        // intercept = ymean - slope * xmean, for y(x) and for x(y); sums by the CTA (float64, fixed order -> deterministic)
        const int n = P.vl_ref > 0 ? min(S.n_rows, P.vl_ref) : S.n_rows;      // VP_REF = 0: all_vp[-0:] is the whole history
        const double *h = rows + 2 * (size_t)(S.n_rows - n);
        const int lane = tid & 31, wid = tid >> 5;
        auto block_sums = [&](double *v, int cnt) {                       // v[0..cnt) -> s_red[q][0] holds the CTA sums
            for (int q = 0; q < cnt; q++) {
                double a = v[q];
                for (int o = 16; o > 0; o >>= 1) a = __dadd_rn(a, __shfl_down_sync(0xffffffffu, a, o));
                if (lane == 0) s_red[q][wid] = a;
            }
            __syncthreads();
            if (tid == 0)
                for (int q = 0; q < cnt; q++) {
                    double a = s_red[q][0];
                    for (int w = 1; w < VP_NT / 32; w++) a = __dadd_rn(a, s_red[q][w]);
                    s_red[q][0] = a;
                }
            __syncthreads();
        };
        double v[3] = {0.0, 0.0, 0.0};
        for (int i = tid; i < n; i += VP_NT) { v[0] = __dadd_rn(v[0], h[2 * i]); v[1] = __dadd_rn(v[1], h[2 * i + 1]); }
        block_sums(v, 2);
        if (tid == 0) { s_mean[0] = __ddiv_rn(s_red[0][0], (double)n); s_mean[1] = __ddiv_rn(s_red[1][0], (double)n); }
        __syncthreads();
        const double xm = s_mean[0], ym = s_mean[1];
        v[0] = v[1] = v[2] = 0.0;
        for (int i = tid; i < n; i += VP_NT) {
            const double dx = __dsub_rn(h[2 * i], xm), dy = __dsub_rn(h[2 * i + 1], ym);
            v[0] = __fma_rn(dx, dx, v[0]); v[1] = __fma_rn(dx, dy, v[1]); v[2] = __fma_rn(dy, dy, v[2]);
        }
        block_sums(v, 3);
        if (tid == 0) {
            const double nn = (double)n;
            const double sxx = __ddiv_rn(s_red[0][0], nn), sxy = __ddiv_rn(s_red[1][0], nn), syy = __ddiv_rn(s_red[2][0], nn);
            const double m = __ddiv_rn(sxy, sxx), mv = __ddiv_rn(sxy, syy);
            S.vl_m = m; S.vl_b = __dsub_rn(ym, __dmul_rn(m, xm));
            S.vl_mv = mv; S.vl_bv = __dsub_rn(xm, __dmul_rn(mv, ym));
            const double bx = S.vp_x, by = S.vp_y, Wm1 = (double)(mW - 1), Hm1 = (double)(mH - 1);
            S.vl_bp[0] = bx; S.vl_bp[1] = by;
            S.vl_lp_y = __dsub_rn(by, __dmul_rn(bx, m));
            S.vl_rp_y = __dadd_rn(by, __dmul_rn(__dsub_rn(Wm1, bx), m));
            S.vl_up_x = __dsub_rn(bx, __dmul_rn(by, mv));
            S.vl_dp_x = __dadd_rn(bx, __dmul_rn(__dsub_rn(Hm1, by), mv));
            S.vl_init = 1;
        }
    }
    if (tid == 0) {
        counts[2 * b] = k0; counts[2 * b + 1] = k1;
        S.refresh = 0;
        if ((double)(k0 + k1) < P.tp_refresh_below || S.tp_ult == P.tp_update_time) { S.tp_ult = 0; S.refresh = 1; }
        if (refresh_out) refresh_out[b] = S.refresh;
        S.tp_ult++;
        S.vp_ult++;
        states[b] = S;
    }
}

// REPlace refresh (VP.py:700-718): masks (0,1) -> group 0, (2,3) -> group 1; only if both groups are non-empty
__global__ void k_apply_refresh(lkf_vp_state *__restrict__ states, const float *__restrict__ corners,
                                const int32_t *__restrict__ ccounts, int max_out, float *__restrict__ p0_all,
                                int32_t *__restrict__ counts, int maxp, int force)
{
    const int b = blockIdx.x;
    if (!force && !states[b].refresh) return;
    // a negative count (-1 - n) reports a capacity overflow of the corner selection: keep the n corners that were written and
    // flag the clip (bit 1 of row_overflow; the host raises when it reads the rows)
    int c[4];
    bool over = false;
    for (int i = 0; i < 4; i++) {
        const int v = ccounts[4 * b + i];
        over |= v < 0;
        c[i] = v < 0 ? -1 - v : v;
    }
    if (over && threadIdx.x == 0) states[b].row_overflow |= 2;
    const int n0 = min(c[0] + c[1], maxp), n1 = min(c[2] + c[3], maxp);
    if (!force && (n0 == 0 || n1 == 0)) return;
    for (int g = 0; g < 2; g++) {
        float *p0 = p0_all + ((size_t)(2 * b + g) * maxp) * 2;
        const int ca = c[2 * g], n = g ? n1 : n0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const float *src = i < ca ? corners + ((size_t)(4 * b + 2 * g) * max_out + i) * 2
                                      : corners + ((size_t)(4 * b + 2 * g + 1) * max_out + (i - ca)) * 2;
            p0[2 * i] = src[0]; p0[2 * i + 1] = src[1];
        }
        if (threadIdx.x == 0) counts[2 * b + g] = n;
    }
}

}  // namespace

LKF_EXPORT int lkf_vp_frame(const lkf_vp_params *P, lkf_vp_state *states, int n_clips, float *p0, const float *p1,
                            const uint8_t *status, int32_t *counts, int max_pts_per_group, const lkf_image *roi_mask,
                            double *rows, int row_capacity, int32_t *refresh_out, const int32_t *clip_enable, lkf_stream stream)
{
    if (!P || !states || n_clips < 1 || !p0 || !p1 || !status || !counts || max_pts_per_group < 1 || !rows || row_capacity < 1)
        return LKF_ERR_ARG;
    if (!img_ok(roi_mask, 1, 1) || roi_mask->batch != 1) return LKF_ERR_ARG;
    if (P->vp_ref_num < 1 || P->vp_ref_num > LKF_VP_RING) return LKF_ERR_UNSUPPORTED;
    const int maxp = max_pts_per_group;
    size_t smem = (size_t)2 * maxp * sizeof(LineRec) + (size_t)8 * maxp * sizeof(float) + (size_t)maxp * sizeof(double) +
                  (size_t)(maxp + 1) * sizeof(int) + (size_t)VP_NT * sizeof(CpRec) + 64;
    if (smem > 200 * 1024) return LKF_ERR_UNSUPPORTED;
    LKF_FUNC_ATTRS(LKF_CUDA(cudaFuncSetAttribute(k_vp_frame, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)));
    k_vp_frame<<<n_clips, VP_NT, smem, (cudaStream_t)stream>>>(*P, states, p0, p1, status, counts, maxp,
                                                               (const uint8_t *)roi_mask->data, roi_mask->row_pitch,
                                                               roi_mask->height, roi_mask->width, rows, row_capacity, refresh_out,
                                                               clip_enable);
    LKF_LAUNCH_CHECK(nullptr);
    return LKF_OK;
}

LKF_EXPORT int lkf_vp_apply_refresh(lkf_vp_state *states, int n_clips, const float *corners, const int32_t *corner_counts,
                                    int max_out, float *p0, int32_t *counts, int max_pts_per_group, int force,
                                    lkf_stream stream)
{
    if (!states || n_clips < 1 || !corners || !corner_counts || !p0 || !counts || max_out < 1 || max_pts_per_group < 1)
        return LKF_ERR_ARG;
    k_apply_refresh<<<n_clips, 64, 0, (cudaStream_t)stream>>>(states, corners, corner_counts, max_out, p0, counts,
                                                             max_pts_per_group, force);
    LKF_LAUNCH_CHECK(nullptr);
    return LKF_OK;
}
