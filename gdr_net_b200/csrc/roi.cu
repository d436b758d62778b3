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

// rot = 0 restatement of get_affine_transform + cv2.warpAffine's inversion: dst = s * (src - center) + out/2, s = out / scale
struct InvAffine {
    double a, bx, by;  // src_x = a * x + bx, src_y = a * y + by
};
__device__ __forceinline__ InvAffine inv_affine(float cx, float cy, float scale, int out) {
    InvAffine m;
    const double s = (double)out / (double)scale;
    m.a = 1.0 / s;
    m.bx = (double)cx - 0.5 * (double)out * m.a;
    m.by = (double)cy - 0.5 * (double)out * m.a;
    return m;
}
constexpr int kAB = 1024;  // AB_SCALE = 1 << 10

// roi_img [B][3][R][R] float = bilinear crop of the uint8 HWC image / 255
__global__ void __launch_bounds__(256) roi_crop_image_kernel(const uint8_t* __restrict__ img, const float* __restrict__ centers,
                                                             const float* __restrict__ scales, float* __restrict__ out, int B, int H,
                                                             int W, int R, float inv_std) {
    const long total = (long)B * R * R;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % R), y = (int)((idx / R) % R), b = (int)(idx / ((long)R * R));
        const InvAffine m = inv_affine(centers[b * 2], centers[b * 2 + 1], scales[b], R);
        const int X = (__double2int_rn(m.bx * kAB) + kAB / 64 + __double2int_rn(m.a * x * kAB)) >> 5;
        const int Y = (__double2int_rn((m.a * y + m.by) * kAB) + kAB / 64) >> 5;
        const int sx = X >> 5, sy = Y >> 5;
        const float ax = (float)(X & 31) * (1.f / 32.f), ay = (float)(Y & 31) * (1.f / 32.f);
        const float w00 = (1.f - ay) * (1.f - ax), w01 = (1.f - ay) * ax, w10 = ay * (1.f - ax), w11 = ay * ax;
        const uint8_t* base = img + (long)b * H * W * 3;
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
                                                          const float* __restrict__ m_trunc, const float* __restrict__ centers,
                                                          const float* __restrict__ scales, const float* __restrict__ extents,
                                                          const float* __restrict__ fps, int n_fps, float* __restrict__ roi_xyz,
                                                          float* __restrict__ o_trunc, float* __restrict__ o_visib, float* __restrict__ o_obj,
                                                          long long* __restrict__ o_region, float* __restrict__ o_coord, int B, int H, int W,
                                                          int R) {
    extern __shared__ float s_fps[];  // [n_fps][3] of this ROI
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < n_fps * 3; i += blockDim.x) s_fps[i] = fps[(long)b * n_fps * 3 + i];
    __syncthreads();
    const InvAffine m = inv_affine(centers[b * 2], centers[b * 2 + 1], scales[b], R);
    const int bxn = __double2int_rn(m.bx * kAB), bxl = bxn;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < R * R; idx += gridDim.x * blockDim.x) {
        const int x = idx % R, y = idx / R;
        const int ad = __double2int_rn(m.a * x * kAB), y0 = __double2int_rn((m.a * y + m.by) * kAB);
        // ---- INTER_NEAREST sample position
        const int Xn = (bxn + kAB / 2 + ad) >> 10, Yn = (y0 + kAB / 2) >> 10;
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
        const int Xl = (bxl + kAB / 64 + ad) >> 5, Yl = (y0 + kAB / 64) >> 5;
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

extern "C" int gdrn_roi_crop_image(const void* image_u8, const float* centers, const float* scales, float* roi_img, int B, int H,
                                   int W, int out_res, float pixel_std, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (B <= 0 || H <= 0 || W <= 0 || out_res <= 0 || pixel_std == 0.f) return set_error(GDRN_ERR_ARG, "roi_crop_image: bad shape");
    const long total = (long)B * out_res * out_res;
    long g = (total + 255) / 256;
    const long cap = (long)num_sms() * 16;
    if (g > cap) g = cap;
    roi_crop_image_kernel<<<(int)g, 256, 0, stream>>>(reinterpret_cast<const uint8_t*>(image_u8), centers, scales, roi_img, B, H, W,
                                                     out_res, 1.f / pixel_std);
    GDRN_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}

extern "C" int gdrn_roi_targets(const float* xyz, const float* mask_visib, const float* mask_trunc, const float* centers,
                                const float* scales, const float* extents, const float* fps_points, int n_fps, float* roi_xyz,
                                float* roi_mask_trunc, float* roi_mask_visib, float* roi_mask_obj, long long* roi_region,
                                float* roi_coord_2d, int B, int H, int W, int out_res, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (B <= 0 || n_fps <= 0 || n_fps > 1024) return set_error(GDRN_ERR_ARG, "roi_targets: bad B=%d / n_fps=%d", B, n_fps);
    dim3 grid((out_res * out_res + 127) / 128, B);
    roi_targets_kernel<<<grid, 128, n_fps * 3 * sizeof(float), stream>>>(xyz, mask_visib, mask_trunc, centers, scales, extents, fps_points,
                                                                        n_fps, roi_xyz, roi_mask_trunc, roi_mask_visib, roi_mask_obj,
                                                                        roi_region, roi_coord_2d, B, H, W, out_res);
    GDRN_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}
