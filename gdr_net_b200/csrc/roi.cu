// GPU crop / target generation of the data loader's per-instance work (reference
// core/gdrn_modeling/data_loader.py:487-560 and core/utils/data_utils.py:80-137, 213-219):
//
//   roi_img        = warpAffine(image, INTER_LINEAR) -> CHW float / 255              (crop_resize_by_warp_affine + normalize_image)
//   roi_coord_2d   = warpAffine(meshgrid linspace(0,1), INTER_LINEAR)                 (2 x 64 x 64)
//   roi_mask_*     = warpAffine(mask, INTER_NEAREST),  mask_obj = [xyz != 0]
//   roi_xyz        = warpAffine(xyz, INTER_NEAREST) / extent + 0.5                     (3 x 64 x 64; every pixel, like the reference)
//   roi_region     = [xyz != 0] * (1 + argmin_f || xyz - fps_f ||)                     (xyz_to_region: scipy cdist + argmin on the host)
//
// The reference does this per instance on dataloader worker CPUs with cv2 / scipy; at > 5 k crops/s per GPU that cannot keep
// up.  The sampling follows cv2.warpAffine's arithmetic exactly: the inverse map is evaluated in 10-bit fixed point
// (X = (round((M01*y + M02)*1024) + round(M00*x*1024) + delta) >> shift), INTER_NEAREST picks pixel (X>>10, Y>>10),
// INTER_LINEAR uses the 1/32-pixel quantised fractions; border pixels are 0 (BORDER_CONSTANT).  float32 sources are
// reproduced to fp32 rounding; the uint8 image differs from cv2 by at most one grey level in < 1 % of the values (cv2
// quantises the 2-D weights to 15-bit integers with a sum fix-up).
#include "gdrn_internal.h"

namespace gdrn {

// The inverse affine map, bit for bit as the reference obtains it:
//   * get_affine_transform (data_utils.py:96-137, rot = 0) builds three source points in FLOAT32 from the float64 centre / scale
//     (src0 = fl32(c), src1 = (fl32(cx), fl32(cy - 0.5 s)), src2 = get_3rd_point in float32) -- these roundings change the scale
//     by ~1e-7 and make exact fixed-point ties COMMON (the point differences are float32 values with few fraction bits);
//   * cv2.getAffineTransform solves the 6 x 6 system with OpenCV's generic LU (partial pivoting, no FMA) in double;
//   * cv2.warpAffine inverts the 2 x 3 matrix in double.
// All three are restated here with explicitly rounded, non-fused operations (__dmul_rn / __dadd_rn): at a tie the rounding
// noise of exactly this operation sequence decides the sampled pixel.  (Checked on the host against cv2.getAffineTransform:
// identical bits in 2000 / 2000 random cases.)
struct InvAffine {
    double m[6];  // src_x = m0 x + m1 y + m2, src_y = m3 x + m4 y + m5
};
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ void solve_inv_affine(double cx, double cy, double scale, int out, InvAffine* res) {
    float src[3][2], dst[3][2];
    src[0][0] = (float)cx;
    src[0][1] = (float)cy;
    src[1][0] = (float)dadd(cx, 0.0);
    src[1][1] = (float)dadd(cy, dmul(scale, -0.5));
    {   // get_3rd_point(a = src0, b = src1) in float32
        const float d0 = __fsub_rn(src[0][0], src[1][0]), d1 = __fsub_rn(src[0][1], src[1][1]);
        src[2][0] = __fadd_rn(src[1][0], -d1);
        src[2][1] = __fadd_rn(src[1][1], d0);
    }
    const float h = 0.5f * (float)out;
    dst[0][0] = h;
    dst[0][1] = h;
    dst[1][0] = h;
    dst[1][1] = __fadd_rn(h, -0.5f * (float)out);
    {
        const float d0 = __fsub_rn(dst[0][0], dst[1][0]), d1 = __fsub_rn(dst[0][1], dst[1][1]);
        dst[2][0] = __fadd_rn(dst[1][0], -d1);
        dst[2][1] = __fadd_rn(dst[1][1], d0);
    }
    // cv2.getAffineTransform: a[6][6] x = b, OpenCV LUImpl (modules/core matrix_decomp: partial pivoting, plain mul / add)
    double A[6][6], b[6];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) A[i][j] = 0.0;
    for (int i = 0; i < 3; ++i) {
        A[2 * i][0] = A[2 * i + 1][3] = (double)src[i][0];
        A[2 * i][1] = A[2 * i + 1][4] = (double)src[i][1];
        A[2 * i][2] = A[2 * i + 1][5] = 1.0;
        b[2 * i] = (double)dst[i][0];
        b[2 * i + 1] = (double)dst[i][1];
    }
    for (int i = 0; i < 6; ++i) {
        int k = i;
        for (int j = i + 1; j < 6; ++j)
            if (fabs(A[j][i]) > fabs(A[k][i])) k = j;
        if (k != i) {
            for (int j = i; j < 6; ++j) {
                const double t = A[i][j];
                A[i][j] = A[k][j];
                A[k][j] = t;
            }
            const double t = b[i];
            b[i] = b[k];
            b[k] = t;
        }
        const double d = -1.0 / A[i][i];
        for (int j = i + 1; j < 6; ++j) {
            const double alpha = dmul(A[j][i], d);
            for (int k2 = i + 1; k2 < 6; ++k2) A[j][k2] = dadd(A[j][k2], dmul(alpha, A[i][k2]));
            b[j] = dadd(b[j], dmul(alpha, b[i]));
        }
    }
    for (int i = 5; i >= 0; --i) {
        double sacc = b[i];
        for (int k2 = i + 1; k2 < 6; ++k2) sacc = dadd(sacc, -dmul(A[i][k2], b[k2]));
        b[i] = sacc / A[i][i];
    }
    // cv2.warpAffine (no WARP_INVERSE_MAP): invert M = [b0 b1 b2; b3 b4 b5]
    double M0 = b[0], M1 = b[1], M2 = b[2], M3 = b[3], M4 = b[4], M5 = b[5];
    double D = dadd(dmul(M0, M4), -dmul(M1, M3));
    D = D != 0.0 ? 1.0 / D : 0.0;
    const double A11 = dmul(M4, D), A22 = dmul(M0, D);
    M0 = A11;
    M1 = dmul(M1, -D);
    M3 = dmul(M3, -D);
    M4 = A22;
    const double b1 = dadd(dmul(-M0, M2), -dmul(M1, M5));
    const double b2 = dadd(dmul(-M3, M2), -dmul(M4, M5));
    res->m[0] = M0;
    res->m[1] = M1;
    res->m[2] = b1;
    res->m[3] = M3;
    res->m[4] = M4;
    res->m[5] = b2;
}
// fixed-point source coordinates of destination pixel (x, y): warpAffine's adelta / bdelta / X0 / Y0 (AB_BITS = 10)
__device__ __forceinline__ void warp_fixed(const InvAffine& a, int x, int y, int round_delta, int& X, int& Y) {
    const int adelta = __double2int_rn(dmul(dmul(a.m[0], (double)x), 1024.0));
    const int bdelta = __double2int_rn(dmul(dmul(a.m[3], (double)x), 1024.0));
    const int X0 = __double2int_rn(dmul(dadd(dmul(a.m[1], (double)y), a.m[2]), 1024.0)) + round_delta;
    const int Y0 = __double2int_rn(dmul(dadd(dmul(a.m[4], (double)y), a.m[5]), 1024.0)) + round_delta;
    X = X0 + adelta;
    Y = Y0 + bdelta;
}
constexpr int kAB = 1024;  // AB_SCALE = 1 << 10

// roi_img [B][3][R][R] float = bilinear crop of the uint8 HWC image / 255; grid (pixel chunks, B)
__global__ void __launch_bounds__(256) roi_crop_image_kernel(const uint8_t* __restrict__ img, const double* __restrict__ centers,
                                                             const double* __restrict__ scales, float* __restrict__ out, int B, int H,
                                                             int W, int R, float inv_std) {
    __shared__ InvAffine m;
    const int b = blockIdx.y;
    if (threadIdx.x == 0) solve_inv_affine(centers[b * 2], centers[b * 2 + 1], scales[b], R, &m);
    __syncthreads();
    const uint8_t* base = img + (long)b * H * W * 3;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < R * R; idx += gridDim.x * blockDim.x) {
        const int x = idx % R, y = idx / R;
        int X, Y;
        warp_fixed(m, x, y, kAB / 64, X, Y);  // round_delta = AB_SCALE / INTER_TAB_SIZE / 2
        X >>= 5;
        Y >>= 5;
        const int sx = X >> 5, sy = Y >> 5;
        const float ax = (float)(X & 31) * (1.f / 32.f), ay = (float)(Y & 31) * (1.f / 32.f);
        const float w00 = (1.f - ay) * (1.f - ax), w01 = (1.f - ay) * ax, w10 = ay * (1.f - ax), w11 = ay * ax;
        float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const int yy = sy + dy, xx = sx + dx;
                if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                    const float w = dy ? (dx ? w11 : w10) : (dx ? w01 : w00);
                    const uint8_t* p = base + ((long)yy * W + xx) * 3;
                    acc[0] = fmaf(w, (float)p[0], acc[0]);
                    acc[1] = fmaf(w, (float)p[1], acc[1]);
                    acc[2] = fmaf(w, (float)p[2], acc[2]);
                }
            }
#pragma unroll
        for (int c = 0; c < 3; ++c)  // cv2 rounds the interpolated value to uint8 before the reference normalises it
            out[(((long)b * 3 + c) * R + y) * R + x] = rintf(acc[c]) * inv_std;
    }
}

// all 64 x 64 targets of one ROI pixel per thread
__global__ void __launch_bounds__(128) roi_targets_kernel(const float* __restrict__ xyz, const float* __restrict__ m_visib,
                                                          const float* __restrict__ m_trunc, const double* __restrict__ centers,
                                                          const double* __restrict__ scales, const float* __restrict__ extents,
                                                          const float* __restrict__ fps, int n_fps, float* __restrict__ roi_xyz,
                                                          float* __restrict__ o_trunc, float* __restrict__ o_visib, float* __restrict__ o_obj,
                                                          long long* __restrict__ o_region, float* __restrict__ o_coord, int B, int H, int W,
                                                          int R) {
    extern __shared__ float s_fps[];  // [n_fps][3] of this ROI
    __shared__ InvAffine m;
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < n_fps * 3; i += blockDim.x) s_fps[i] = fps[(long)b * n_fps * 3 + i];
    if (threadIdx.x == 0) solve_inv_affine(centers[b * 2], centers[b * 2 + 1], scales[b], R, &m);
    __syncthreads();
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < R * R; idx += gridDim.x * blockDim.x) {
        const int x = idx % R, y = idx / R;
        // ---- INTER_NEAREST sample position (round_delta = AB_SCALE / 2)
        int Xn, Yn;
        warp_fixed(m, x, y, kAB / 2, Xn, Yn);
        Xn >>= 10;
        Yn >>= 10;
        const bool in = Xn >= 0 && Xn < W && Yn >= 0 && Yn < H;
        const long src = ((long)b * H + Yn) * W + Xn;
        float v[3] = {0.f, 0.f, 0.f}, mv = 0.f, mt = 0.f, mo = 0.f;
        if (in) {
            v[0] = xyz[src * 3];
            v[1] = xyz[src * 3 + 1];
            v[2] = xyz[src * 3 + 2];
            // full-resolution masks (data_loader.py:472, 507-512): obj = [xyz != 0], visib = seg * obj, trunc = visib * trunc
            const long s2 = src;
            const float obj = (v[0] != 0.f || v[1] != 0.f || v[2] != 0.f) ? 1.f : 0.f;
            mo = obj;
            mv = m_visib[s2] * obj;
            mt = m_trunc != nullptr ? mv * m_trunc[s2] : mv;
        }
        const long o = (long)b * R * R + idx;
        o_obj[o] = mo;
        o_visib[o] = mv;
        o_trunc[o] = mt;
        // ---- region label from the UN-normalised crop (xyz_to_region: double-precision distances, first minimum)
        const bool fg = (v[0] != 0.f || v[1] != 0.f || v[2] != 0.f);
        int best = 0;
        double bd = 1e300;
        for (int f = 0; f < n_fps; ++f) {
            const double dx = (double)v[0] - s_fps[f * 3], dy = (double)v[1] - s_fps[f * 3 + 1], dz = (double)v[2] - s_fps[f * 3 + 2];
            const double d = sqrt(dx * dx + dy * dy + dz * dz);
            if (d < bd) {
                bd = d;
                best = f;
            }
        }
        o_region[o] = fg ? (long long)(best + 1) : 0;
        // ---- xyz normalised by the extent, every pixel (data_loader.py:538-542)
#pragma unroll
        for (int c = 0; c < 3; ++c) roi_xyz[((long)b * 3 + c) * R * R + idx] = v[c] / extents[b * 3 + c] + 0.5f;
        // ---- roi_coord_2d: INTER_LINEAR over the meshgrid linspace(0,1,W) x linspace(0,1,H) (border 0)
        int Xl, Yl;
        warp_fixed(m, x, y, kAB / 64, Xl, Yl);
        Xl >>= 5;
        Yl >>= 5;
        const int sx = Xl >> 5, sy = Yl >> 5;
        const float ax = (float)(Xl & 31) * (1.f / 32.f), ay = (float)(Yl & 31) * (1.f / 32.f);
        float cx = 0.f, cy = 0.f;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const int yy = sy + dy, xx = sx + dx;
                if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                    const float w = (dy ? ay : 1.f - ay) * (dx ? ax : 1.f - ax);
                    // np.linspace(0, 1, n, dtype=float32)[i]: i * (1 / (n - 1)) evaluated in double, then rounded to float32
                    cx = fmaf(w, (float)((double)xx * (1.0 / (double)(W - 1))), cx);
                    cy = fmaf(w, (float)((double)yy * (1.0 / (double)(H - 1))), cy);
                }
            }
        o_coord[((long)b * 2 + 0) * R * R + idx] = cx;
        o_coord[((long)b * 2 + 1) * R * R + idx] = cy;
    }
}

}  // namespace gdrn

using namespace gdrn;

extern "C" int gdrn_roi_crop_image(const void* image_u8, const double* centers, const double* scales, float* roi_img, int B, int H,
                                   int W, int out_res, float pixel_std, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (B <= 0 || H <= 0 || W <= 0 || out_res <= 0 || pixel_std == 0.f) return set_error(GDRN_ERR_ARG, "roi_crop_image: bad shape");
    int chunks = (out_res * out_res + 255) / 256;
    if (chunks > 64) chunks = 64;
    GDRN_LAUNCH_SMEM(roi_crop_image_kernel, dim3(chunks, B), dim3(256), 0, stream, reinterpret_cast<const uint8_t*>(image_u8), centers, scales,
                     roi_img, B, H, W, out_res, 1.f / pixel_std);
    GDRN_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}

extern "C" int gdrn_roi_targets(const float* xyz, const float* mask_visib, const float* mask_trunc, const double* centers,
                                const double* scales, const float* extents, const float* fps_points, int n_fps, float* roi_xyz,
                                float* roi_mask_trunc, float* roi_mask_visib, float* roi_mask_obj, long long* roi_region,
                                float* roi_coord_2d, int B, int H, int W, int out_res, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (B <= 0 || n_fps <= 0 || n_fps > 1024) return set_error(GDRN_ERR_ARG, "roi_targets: bad B=%d / n_fps=%d", B, n_fps);
    dim3 grid((out_res * out_res + 127) / 128, B);
    GDRN_LAUNCH_SMEM(roi_targets_kernel, grid, dim3(128), n_fps * 3 * sizeof(float), stream, xyz, mask_visib, mask_trunc, centers, scales,
                     extents, fps_points, n_fps, roi_xyz, roi_mask_trunc, roi_mask_visib, roi_mask_obj, roi_region, roi_coord_2d, B, H, W,
                     out_res);
    GDRN_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}
