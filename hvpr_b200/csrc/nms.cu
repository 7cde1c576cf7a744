// N3 — single-stage post-processing: Detector3DTemplate.post_processing (pcdet/models/detectors/
// detector3d_template.py:168-260, the class-agnostic branch) + class_agnostic_nms (pcdet/models/model_utils/model_nms_utils.py:6-25):
//   sigmoid, max over classes -> score threshold -> top NMS_PRE_MAXSIZE by score -> rotated-BEV-IoU NMS -> first NMS_POST_MAXSIZE.
// The reference's NMS itself is the `iou3d_nms` CUDA op, declared in setup.py:53-62 but NOT in the tree: PARITY UNPINNED for the IoU
// arithmetic — this file restates the published algorithm (greedy NMS in descending score order on the rotated bird's-eye-view IoU of
// (x, y, dx, dy, heading) rectangles); oracle/post_process.py is an independent float64 restatement.
//
//   pp_score_kernel    one thread per anchor: score/threshold, append a 64-bit key (score bits : ~index) to the frame's candidate list
//   pp_topk_kernel     one block per frame: exact radix select of the K largest keys (6 x 11-bit passes), then a bitonic sort in shared
//                      memory -> candidates in descending score order (ties: lower anchor index first), deterministic
//   pp_iou_mask_kernel 64 x 64 tiles of the upper triangle: intersection area of two rotated rectangles as a boundary integral over
//                      Liang-Barsky-clipped edges (registers only), bit j of mask[i] = IoU > thresh
//   pp_reduce_kernel   one block per frame: 64 boxes at a time, the diagonal word resolved serially, kept rows OR-ed into the
//                      suppression words; gathers the kept boxes / scores / labels
#include "common.cuh"
#include <math.h>

namespace hvpr {

constexpr int kPpMaxPre = 4096;        // NMS_PRE_MAXSIZE capacity (multiple of 64, power of two for the bitonic sort)
constexpr int kPpWords = kPpMaxPre / 64;

struct PpWorkspace {
    unsigned long long *cand;     // [B][N] candidate keys (unordered)
    int32_t *count;               // [B]
    unsigned long long *sorted;   // [B][kPpMaxPre] keys in descending order
    int32_t *nsel;                // [B] number of valid entries in `sorted`
    unsigned long long *mask;     // [B][kPpMaxPre][kPpWords]
    size_t bytes;
};
static PpWorkspace pp_carve(void *base, int B, int64_t N) {
    PpWorkspace w;
    size_t off = 0;
    auto take = [&](size_t bytes) { void *p = base ? (char *)base + off : nullptr; off += align_up(bytes, 256); return p; };
    w.cand = (unsigned long long *)take(sizeof(unsigned long long) * (size_t)B * N);
    w.count = (int32_t *)take(sizeof(int32_t) * B);
    w.sorted = (unsigned long long *)take(sizeof(unsigned long long) * (size_t)B * kPpMaxPre);
    w.nsel = (int32_t *)take(sizeof(int32_t) * B);
    w.mask = (unsigned long long *)take(sizeof(unsigned long long) * (size_t)B * kPpMaxPre * kPpWords);
    w.bytes = off;
    return w;
}

__device__ __forceinline__ float pp_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---- score + threshold ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pp_score_kernel(const float *__restrict__ cls, int64_t N, int C, int normalized, float thresh,
                                                       unsigned long long *__restrict__ cand, int32_t *__restrict__ count) {
    const int f = blockIdx.y;
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    bool pass = false;
    float s = 0.0f;
    if (i < N) {
        const float *p = cls + ((int64_t)f * N + i) * C;
        float best = p[0];
        for (int c = 1; c < C; ++c) best = fmaxf(best, p[c]);
        s = normalized ? best : pp_sigmoid(best);          // sigmoid is monotone: max of sigmoids = sigmoid of max (:207, :241)
        pass = s >= thresh;
    }
    const unsigned m = __ballot_sync(0xffffffffu, pass);
    if (m == 0) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0) base = atomicAdd(count + f, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (pass) {
        const unsigned long long key = ((unsigned long long)__float_as_uint(s) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)i);
        cand[(int64_t)f * N + base + __popc(m & ((1u << lane) - 1u))] = key;
    }
}

// ---- exact top-K of the candidate keys, sorted descending ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) pp_topk_kernel(const unsigned long long *__restrict__ cand, const int32_t *__restrict__ count,
                                                       int64_t N, int K, unsigned long long *__restrict__ sorted, int32_t *__restrict__ nsel) {
    __shared__ unsigned long long keys[kPpMaxPre];
    __shared__ int hist[2048];
    __shared__ unsigned long long s_prefix;
    __shared__ int s_remaining, s_fill;
    __shared__ int s_wsum[32];
    const int f = blockIdx.x, t = threadIdx.x;
    const unsigned long long *src = cand + (int64_t)f * N;
    const int n = count[f];
    const int k = n < K ? n : K;
    for (int i = t; i < kPpMaxPre; i += 1024) keys[i] = 0ull;
    if (t == 0) { s_prefix = 0ull; s_remaining = k; s_fill = 0; }
    __syncthreads();
    unsigned long long thr = 0ull;                      // keys >= thr are selected
    if (n > K) {
        // radix select, most significant digits first: after the passes s_prefix is the K-th largest key (keys are unique)
        for (int shift = 55; shift >= -11; shift -= 11) {          // digits at bit offsets 55, 44, 33, 22, 11, 0 (the top digit is 9 bits)
            const int sh = shift < 0 ? 0 : shift;
            const int bits = shift == 55 ? 9 : 11;
            const unsigned long long hi_mask = (sh + bits >= 64) ? 0ull : (~0ull << (sh + bits));
            for (int i = t; i < 2048; i += 1024) hist[i] = 0;
            __syncthreads();
            const unsigned long long prefix = s_prefix;
            for (int i0 = 0; i0 < n; i0 += 1024) {             // whole block iterates together: __match_any_sync needs full warps
                const int i = i0 + t;
                int digit = -1;                                 // -1: not a contender (out of range or different prefix)
                if (i < n) {
                    const unsigned long long key = src[i];
                    if ((key & hi_mask) == (prefix & hi_mask)) digit = (int)((key >> sh) & ((1u << bits) - 1u));
                }
                // sigmoid scores share their leading digits: aggregate equal digits inside the warp, one shared-memory atomic per value
                const unsigned peers = __match_any_sync(0xffffffffu, digit);
                if (digit >= 0 && (t & 31) == __ffs(peers) - 1) atomicAdd(&hist[digit], __popc(peers));
            }
            __syncthreads();
            // find the digit d with  sum(hist[d+1..]) < remaining <= sum(hist[d..])  in parallel: thread t owns bins 2t, 2t+1, a
            // suffix scan inside each warp, then over the 32 warp totals (a serial scan of 2048 bins by one thread was 80 % of the kernel)
            {
                const int lane = t & 31, wid = t >> 5;
                const int h0 = hist[2 * t], h1 = hist[2 * t + 1];
                int suf = h0 + h1;                                   // inclusive suffix sum over the warp's lanes (higher bins first)
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_down_sync(0xffffffffu, suf, o); if (lane + o < 32) suf += v; }
                if (lane == 0) s_wsum[wid] = suf;                    // total of this warp's 64 bins
                const int rem = s_remaining;                         // read before the barrier: one thread rewrites it below
                __syncthreads();
                int above = 0;                                       // bins of all higher warps
                for (int w2 = wid + 1; w2 < 32; ++w2) above += s_wsum[w2];
                const int incl = above + suf;                        // sum over bins >= 2t
                const int excl_pair = incl - (h0 + h1);              // sum over bins >= 2t+2
                if (excl_pair < rem && rem <= incl) {                // the target digit is 2t+1 or 2t: exactly one thread gets here
                    const int d = (excl_pair + h1 >= rem) ? 2 * t + 1 : 2 * t;
                    s_prefix = prefix | ((unsigned long long)d << sh);
                    s_remaining = rem - (d == 2 * t + 1 ? excl_pair : excl_pair + h1);
                }
            }
            __syncthreads();
            if (sh == 0) break;
        }
        thr = s_prefix;
    }
    for (int i = t; i < n; i += 1024) {
        const unsigned long long key = src[i];
        if (key >= thr) { const int p = atomicAdd(&s_fill, 1); if (p < kPpMaxPre) keys[p] = key; }
    }
    __syncthreads();
    // bitonic sort, descending (empty slots are 0 and sink to the end)
    for (int kk = 2; kk <= kPpMaxPre; kk <<= 1)
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int i = t; i < kPpMaxPre; i += 1024) {
                const int l = i ^ j;
                if (l > i) {
                    const unsigned long long a = keys[i], b = keys[l];
                    const bool desc = (i & kk) == 0;
                    if (desc ? (a < b) : (a > b)) { keys[i] = b; keys[l] = a; }
                }
            }
            __syncthreads();
        }
    for (int i = t; i < kPpMaxPre; i += 1024) sorted[(int64_t)f * kPpMaxPre + i] = keys[i];
    if (t == 0) nsel[f] = k;
}

// ---- rotated BEV IoU ------------------------------------------------------------------------------------------------------------------
struct PpRect { float cx, cy, hx, hy, c, s, area, rad2; };      // centre, half sizes, cos/sin of the heading, squared circumradius
__device__ __forceinline__ PpRect pp_rect(const float *b) {
    PpRect r;
    r.cx = b[0]; r.cy = b[1]; r.hx = 0.5f * b[3]; r.hy = 0.5f * b[4];
    sincosf(b[6], &r.s, &r.c);
    r.area = b[3] * b[4];
    r.rad2 = r.hx * r.hx + r.hy * r.hy;
    return r;
}
// Area of (rectangle a) ∩ (rectangle b) without building the intersection polygon: by Green's theorem the area is the boundary integral
// 1/2 ∮ (x dy - y dx), and the boundary of the intersection consists of the parts of a's edges inside b plus the parts of b's edges
// inside a (both counter-clockwise).  Each edge is clipped against the other rectangle's slabs (Liang-Barsky) in that rectangle's
// local frame, where they are axis-aligned — eight segments, fully unrolled, registers only.
//
// Coincident / collinear edges (exact duplicates, heading + pi duplicates, same-heading boxes sharing an edge line — the boxes NMS
// exists to remove) must be counted exactly ONCE.  A closed-vs-open test on `dx == 0` cannot do that in fp32: after the rotate /
// translate round trip a coincident edge sits ~1e-6 m off the slab face and lands inside, outside or astride it at random (round 1:
// inter/area spread over [0, 2] for bit-identical boxes).  Instead the two clip regions are made decisively different:
//     a's edges are clipped against b GROWN by kPpEps   (an edge of a along b's boundary is inside  -> counted),
//     b's edges are clipped against a SHRUNK by kPpEps  (the same edge of b is outside               -> not counted),
// and every corner is formed as (centre difference) + (local offset), so the coordinate noise is an ulp of a few metres (~5e-7 m),
// far below kPpEps.  The price is an area error of at most kPpEps x perimeter (~1e-4 m^2 on a 6 m^2 car box: 2e-5 relative).
constexpr float kPpEps = 1e-5f;
// segment p(t) = p0 + t (p1 - p0), t in [0,1], clipped to |x| <= hx, |y| <= hy: the surviving parameter range [t0, t1] (empty: t0 >= t1)
__device__ __forceinline__ bool pp_clip(float x0, float y0, float dx, float dy, float hx, float hy, float &t0, float &t1) {
    t0 = 0.0f; t1 = 1.0f;
    bool ok = true;
    // |d| below 1e-12: the edge is parallel to the slab (a reciprocal would overflow); one reciprocal per axis otherwise
    if (fabsf(dx) < 1e-12f) ok = fabsf(x0) <= hx;
    else { const float inv = 1.0f / dx; const float ra = (hx - x0) * inv, rb = (-hx - x0) * inv; t1 = fminf(t1, fmaxf(ra, rb)); t0 = fmaxf(t0, fminf(ra, rb)); }
    if (fabsf(dy) < 1e-12f) ok = ok && fabsf(y0) <= hy;
    else { const float inv = 1.0f / dy; const float ra = (hy - y0) * inv, rb = (-hy - y0) * inv; t1 = fminf(t1, fmaxf(ra, rb)); t0 = fmaxf(t0, fminf(ra, rb)); }
    return ok && t0 < t1;
}
__device__ __forceinline__ float pp_intersection(const PpRect &a, const PpRect &b) {
    // Every piece is evaluated in ONE coordinate system (b's local frame): the closed integral is origin-independent, its parts are not.
    const float lxa[4] = {a.hx, -a.hx, -a.hx, a.hx}, lya[4] = {a.hy, a.hy, -a.hy, -a.hy};
    const float lxb[4] = {b.hx, -b.hx, -b.hx, b.hx}, lyb[4] = {b.hy, b.hy, -b.hy, -b.hy};
    const float ddx = a.cx - b.cx, ddy = a.cy - b.cy;          // exact or nearly so: the centres are close whenever the boxes can overlap
    float px[4], py[4], qx[4], qy[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        // a's corners (counter-clockwise) in b's frame
        const float wx = ddx + (lxa[i] * a.c - lya[i] * a.s), wy = ddy + (lxa[i] * a.s + lya[i] * a.c);
        px[i] = wx * b.c + wy * b.s;
        py[i] = -wx * b.s + wy * b.c;
        // b's corners in a's frame
        const float vx = (lxb[i] * b.c - lyb[i] * b.s) - ddx, vy = (lxb[i] * b.s + lyb[i] * b.c) - ddy;
        qx[i] = vx * a.c + vy * a.s;
        qy[i] = -vx * a.s + vy * a.c;
    }
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int j = (i + 1) & 3;
        float t0, t1;
        // a's edge i inside b grown by eps, evaluated where it lies (b's frame)
        {
            const float dx = px[j] - px[i], dy = py[j] - py[i];
            if (pp_clip(px[i], py[i], dx, dy, b.hx + kPpEps, b.hy + kPpEps, t0, t1)) {
                const float ax = px[i] + t0 * dx, ay = py[i] + t0 * dy, bx = px[i] + t1 * dx, by = py[i] + t1 * dy;
                acc += 0.5f * (ax * by - bx * ay);
            }
        }
        // b's edge i inside a shrunk by eps: clip parameters from a's frame (they are frame-independent), positions taken in b's
        // frame, where b's corners are simply (+-hx, +-hy)
        {
            const float dx = qx[j] - qx[i], dy = qy[j] - qy[i];
            if (pp_clip(qx[i], qy[i], dx, dy, a.hx - kPpEps, a.hy - kPpEps, t0, t1)) {
                const float ex = lxb[j] - lxb[i], ey = lyb[j] - lyb[i];
                const float ax = lxb[i] + t0 * ex, ay = lyb[i] + t0 * ey, bx = lxb[i] + t1 * ex, by = lyb[i] + t1 * ey;
                acc += 0.5f * (ax * by - bx * ay);
            }
        }
    }
    return fminf(fmaxf(acc, 0.0f), fminf(a.area, b.area));
}

__global__ void __launch_bounds__(64) pp_iou_mask_kernel(const float *__restrict__ boxes, int64_t N, const unsigned long long *__restrict__ sorted,
                                                         const int32_t *__restrict__ nsel, float thresh, unsigned long long *__restrict__ mask) {
    const int f = blockIdx.z, rb = blockIdx.y, cb = blockIdx.x;
    if (cb < rb) return;                                  // upper triangle: a box only suppresses lower-scored ones
    const int n = nsel[f];
    if (rb * 64 >= n || cb * 64 >= n) return;
    __shared__ PpRect crect[64];                          // column rectangles, sin/cos evaluated once per box (not once per pair)
    const int t = threadIdx.x;
    const unsigned long long *keys = sorted + (int64_t)f * kPpMaxPre;
    const int cj = cb * 64 + t;
    if (cj < n) {
        const int64_t idx = (int64_t)(0xFFFFFFFFu - (uint32_t)(keys[cj] & 0xFFFFFFFFull));
        crect[t] = pp_rect(boxes + ((int64_t)f * N + idx) * 7);
    }
    __syncthreads();
    const int ri = rb * 64 + t;
    if (ri >= n) return;
    const PpRect a = (rb == cb) ? crect[t] : pp_rect(boxes + ((int64_t)f * N + (int64_t)(0xFFFFFFFFu - (uint32_t)(keys[ri] & 0xFFFFFFFFull))) * 7);
    unsigned long long bits = 0ull;
    const int cols = (n - cb * 64) < 64 ? (n - cb * 64) : 64;
    for (int j = (rb == cb) ? t + 1 : 0; j < cols; ++j) {
        const PpRect b = crect[j];
        // cheap reject: circumscribed circles do not touch -> no overlap (most pairs of a 70 m x 80 m scene)
        const float ddx = a.cx - b.cx, ddy = a.cy - b.cy;
        const float rsum2 = a.rad2 + b.rad2 + 2.0f * sqrtf(a.rad2 * b.rad2);
        if (ddx * ddx + ddy * ddy > rsum2) continue;
        const float inter = pp_intersection(a, b);
        const float iou = inter / fmaxf(a.area + b.area - inter, 1e-8f);
        if (iou > thresh) bits |= 1ull << j;
    }
    mask[((int64_t)f * kPpMaxPre + ri) * kPpWords + cb] = bits;
}

// ---- greedy reduction + gather ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) pp_reduce_kernel(const float *__restrict__ boxes, const float *__restrict__ cls, int64_t N, int C,
                                                       const unsigned long long *__restrict__ sorted, const int32_t *__restrict__ nsel,
                                                       const unsigned long long *__restrict__ mask, int post_max,
                                                       float *__restrict__ out_boxes, float *__restrict__ out_scores,
                                                       int32_t *__restrict__ out_labels, int32_t *__restrict__ out_index,
                                                       int32_t *__restrict__ out_count) {
    const int f = blockIdx.x, t = threadIdx.x;
    const int n = nsel[f];
    const int nblk = (n + 63) / 64;
    __shared__ unsigned long long s_keep;
    __shared__ int s_kept;
    __shared__ int s_list[64];
    unsigned long long removed = 0ull;                   // suppression word t (boxes 64t .. 64t+63)
    const unsigned long long *keys = sorted + (int64_t)f * kPpMaxPre;
    const unsigned long long *mk = mask + (int64_t)f * kPpMaxPre * kPpWords;
    if (t == 0) s_kept = 0;
    __syncthreads();
    __shared__ unsigned long long s_diag[64];
    for (int blk = 0; blk < nblk; ++blk) {
        // the 64 x 64 diagonal tile of this block: one coalesced-ish load per thread instead of 64 dependent global loads below
        s_diag[t] = (blk * 64 + t < n) ? mk[(int64_t)(blk * 64 + t) * kPpWords + blk] : 0ull;
        __syncthreads();
        if (t == blk) {
            // resolve the diagonal word serially: box j survives iff no kept box of higher score suppresses it
            unsigned long long alive = ~removed, keep = 0ull;
            const int lim = (n - blk * 64) < 64 ? (n - blk * 64) : 64;
            int kept = s_kept, m = 0;
            for (int j = 0; j < lim && kept < post_max; ++j)
                if ((alive >> j) & 1ull) {
                    keep |= 1ull << j;
                    s_list[m++] = j;
                    ++kept;
                    alive &= ~s_diag[j];
                }
            s_keep = keep;
        }
        __syncthreads();
        const unsigned long long keep = s_keep;
        const int base = s_kept;
        const int m = __popcll(keep);
        // gather the kept boxes of this block (thread j < m writes one box) and OR their rows into the later suppression words
        if (t < m) {
            const int j = s_list[t];
            const unsigned long long key = keys[blk * 64 + j];
            const int64_t idx = (int64_t)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
            const int o = base + t;
            for (int e = 0; e < 7; ++e) out_boxes[((int64_t)f * post_max + o) * 7 + e] = boxes[((int64_t)f * N + idx) * 7 + e];
            out_scores[(int64_t)f * post_max + o] = __uint_as_float((uint32_t)(key >> 32));
            const float *p = cls + ((int64_t)f * N + idx) * C;
            int lab = 0;
            for (int c = 1; c < C; ++c) if (p[c] > p[lab]) lab = c;
            out_labels[(int64_t)f * post_max + o] = lab + 1;                 // :241-246 label_preds + 1
            out_index[(int64_t)f * post_max + o] = (int32_t)idx;
        }
        if (t > blk && t < nblk) {
            // rows of the kept boxes, eight loads in flight at a time (a load-then-OR loop serialises one L2 round trip per kept box)
            for (int q0 = 0; q0 < m; q0 += 8) {
                unsigned long long rowsv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    rowsv[u] = (q0 + u < m) ? mk[(int64_t)(blk * 64 + s_list[q0 + u]) * kPpWords + t] : 0ull;
#pragma unroll
                for (int u = 0; u < 8; ++u) removed |= rowsv[u];
            }
        }
        __syncthreads();
        if (t == 0) s_kept = base + m;
        __syncthreads();
        if (s_kept >= post_max) break;
    }
    if (t == 0) out_count[f] = s_kept;
}

// pairwise 3-D IoU of upright boxes [x, y, z, dx, dy, dz, heading] (z = box centre): rotated-BEV overlap x height overlap over the
// union volume — the published boxes_iou3d_gpu the reference calls from generate_recall_record (detector3d_template.py:277-310)
__global__ void __launch_bounds__(256) pp_iou3d_kernel(const float *__restrict__ a, int n, const float *__restrict__ b, int m,
                                                       float *__restrict__ iou) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= (int64_t)n * m) return;
    const float *pa = a + (i / m) * 7, *pb = b + (i % m) * 7;
    const PpRect ra = pp_rect(pa), rb = pp_rect(pb);
    float inter = 0.0f;
    const float ddx = ra.cx - rb.cx, ddy = ra.cy - rb.cy;
    const float rsum2 = ra.rad2 + rb.rad2 + 2.0f * sqrtf(ra.rad2 * rb.rad2);
    if (ddx * ddx + ddy * ddy <= rsum2) inter = pp_intersection(ra, rb);
    const float top = fminf(pa[2] + 0.5f * pa[5], pb[2] + 0.5f * pb[5]), bot = fmaxf(pa[2] - 0.5f * pa[5], pb[2] - 0.5f * pb[5]);
    const float o3d = inter * fmaxf(top - bot, 0.0f);
    const float va = pa[3] * pa[4] * pa[5], vb = pb[3] * pb[4] * pb[5];
    iou[i] = o3d / fmaxf(va + vb - o3d, 1e-6f);
}

}  // namespace hvpr
using namespace hvpr;

extern "C" int hvpr_boxes_iou3d(const float *boxes_a, int n, const float *boxes_b, int m, float *iou, void *stream_) {
    if (n < 0 || m < 0) return HVPR_ERR_ARG;
    if (n == 0 || m == 0) return HVPR_OK;
    if (!boxes_a || !boxes_b || !iou) return HVPR_ERR_ARG;
    pp_iou3d_kernel<<<(unsigned)ceil_div64((int64_t)n * m, 256), 256, 0, (cudaStream_t)stream_>>>(boxes_a, n, boxes_b, m, iou);
    HVPR_CHECK_LAUNCH();
    return HVPR_OK;
}

extern "C" size_t hvpr_post_process_workspace_bytes(int n_frames, int64_t n_boxes) {
    if (n_frames <= 0 || n_boxes <= 0) return 0;
    return pp_carve(nullptr, n_frames, n_boxes).bytes;
}

extern "C" int hvpr_post_process(const float *cls_preds, const float *box_preds, int n_frames, int64_t n_boxes, int num_class,
                                 int cls_normalized, float score_thresh, int nms_pre_max, int nms_post_max, float nms_thresh,
                                 float *out_boxes, float *out_scores, int32_t *out_labels, int32_t *out_index, int32_t *out_count,
                                 void *workspace, size_t workspace_bytes, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!cls_preds || !box_preds || !out_boxes || !out_scores || !out_labels || !out_index || !out_count) return HVPR_ERR_ARG;
    if (n_frames <= 0 || n_boxes <= 0 || num_class <= 0 || nms_post_max <= 0 || n_boxes >= 0x7FFFFFFFll) return HVPR_ERR_ARG;
    if (nms_pre_max <= 0 || nms_pre_max > kPpMaxPre) return HVPR_ERR_UNSUPPORTED;
    PpWorkspace w = pp_carve(workspace, n_frames, n_boxes);
    if (!workspace || workspace_bytes < w.bytes) return HVPR_ERR_WORKSPACE;
    HVPR_CHECK_CUDA(cudaMemsetAsync(w.count, 0, sizeof(int32_t) * n_frames, stream));
    pp_score_kernel<<<dim3((unsigned)ceil_div64(n_boxes, 256), (unsigned)n_frames), 256, 0, stream>>>(
        cls_preds, n_boxes, num_class, cls_normalized, score_thresh, w.cand, w.count);
    HVPR_CHECK_LAUNCH();
    pp_topk_kernel<<<n_frames, 1024, 0, stream>>>(w.cand, w.count, n_boxes, nms_pre_max, w.sorted, w.nsel);
    HVPR_CHECK_LAUNCH();
    pp_iou_mask_kernel<<<dim3(kPpWords, kPpWords, (unsigned)n_frames), 64, 0, stream>>>(box_preds, n_boxes, w.sorted, w.nsel, nms_thresh, w.mask);
    HVPR_CHECK_LAUNCH();
    pp_reduce_kernel<<<n_frames, 64, 0, stream>>>(box_preds, cls_preds, n_boxes, num_class, w.sorted, w.nsel, w.mask, nms_post_max,
                                                  out_boxes, out_scores, out_labels, out_index, out_count);
    HVPR_CHECK_LAUNCH();
    return HVPR_OK;
}
