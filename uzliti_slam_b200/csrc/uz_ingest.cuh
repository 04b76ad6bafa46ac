// uz_ingest.cuh — keyframe ingestion on the device: the two data formats in front of the store.
//
//   backproject_kernel      FeatureExtractionCore::extract3dFeatures
//                           (/root/reference/feature_extraction/src/feature_extraction_core.cpp:254-295): pixel (u, v)
//                           clamped into the depth image, depth read as float -> double, valid iff depth != 0, not NaN
//                           and (max_depth == 0 or depth <= max_depth); x = (u - cx) * depth / fx, y likewise, z = depth;
//                           otherwise (0, 0, -1) and is_3d = false.  All in double, no FMA (the expression is mul then div).
//   wire_decode_kernel      FeatureData::fromMsg (/root/reference/graph_slam_common/src/sensor_data.cpp:124-171) applied to
//                           the ROS1-serialised graph_slam_msgs/Feature[] (graph_slam_msgs/msg/Feature.msg:1-13): per
//                           element  int32 u, int32 v, uint8 is_3d, float32 keypoint_strength, uint32 len, len x float32
//                           descriptor, float64 x, y, z  — little endian, unpadded, so a 32-column element is 169 bytes
//                           and a 64-column one (BRISK, FREAK) 297.
//                           Descriptor values are narrowed float -> unsigned char exactly as the x86 build does
//                           ((unsigned char)val == low byte of cvttss2si).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace uz {

struct CameraModel { double fx, fy, cx, cy, max_depth; int32_t width, height; };

// out index: reverse ? n-1-i : i  (the reference walks features_2d back to front and push_back()s, :263)
__global__ void __launch_bounds__(256) backproject_kernel(const int32_t* __restrict__ u, const int32_t* __restrict__ v, int n,
                                                          const float* __restrict__ depth, int depth_stride_floats, CameraModel cam,
                                                          int reverse, double* __restrict__ pos, uint8_t* __restrict__ valid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int uu = u[i], vv = v[i];                           // Feature.u/v are int32: round() is the identity (:264,:270)
    uu = uu < 0 ? 0 : (uu >= cam.width ? cam.width - 1 : uu);
    vv = vv < 0 ? 0 : (vv >= cam.height ? cam.height - 1 : vv);
    const double d = (double)depth[(size_t)vv * depth_stride_floats + uu];
    const int o = reverse ? n - 1 - i : i;
    double x = 0.0, y = 0.0, z = -1.0;
    uint8_t ok = 0;
    if (d != 0.0 && !(d != d) && (cam.max_depth == 0.0 || d <= cam.max_depth)) {
        z = d;
        x = __ddiv_rn(__dmul_rn(__dsub_rn((double)uu, cam.cx), d), cam.fx);
        y = __ddiv_rn(__dmul_rn(__dsub_rn((double)vv, cam.cy), d), cam.fy);
        ok = 1;
    }
    pos[3 * (size_t)o] = x; pos[3 * (size_t)o + 1] = y; pos[3 * (size_t)o + 2] = z;
    valid[o] = ok;
}

// rows of a descriptor matrix in reverse order (companion of backproject's reverse mode)
__global__ void __launch_bounds__(256) reverse_rows_kernel(const uint8_t* __restrict__ src, int n, int stride, int row_words,
                                                           uint32_t* __restrict__ dst) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;        // one thread per 4 bytes
    if (e >= n * row_words) return;
    const int i = e / row_words, w = e - i * row_words;
    const uint8_t* p = src + (size_t)(n - 1 - i) * stride + 4 * w;
    dst[e] = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

__device__ __forceinline__ uint32_t load_u32_unaligned(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
__device__ __forceinline__ double load_f64_unaligned(const uint8_t* p) {
    const unsigned long long lo = load_u32_unaligned(p), hi = load_u32_unaligned(p + 4);
    return __longlong_as_double((long long)(lo | (hi << 32)));
}
// (unsigned char)val as gcc/x86 evaluates it: cvttss2si (0x80000000 on overflow/NaN), low byte
__device__ __forceinline__ uint8_t narrow_x86(float val) {
    if (!(val > -2147483904.0f && val < 2147483648.0f)) return 0;
    return (uint8_t)(__float2int_rz(val) & 0xFF);
}

__host__ __device__ constexpr int wire_elem_bytes(int cols) { return 4 + 4 + 1 + 4 + 4 + cols * 4 + 24; }      // 169 / 297
static_assert(wire_elem_bytes(32) == 169 && wire_elem_bytes(64) == 297, "Feature.msg element size");

// blob points at the first element (behind the uint32 element count).  One warp per feature, cols = 32 or 64 columns
// (the first element's length, read by the host).  status[0] is set to 1 if an element's descriptor length differs.
__global__ void __launch_bounds__(256) wire_decode_kernel(const uint8_t* __restrict__ blob, int n, int cols, uint8_t* __restrict__ desc,
                                                          double* __restrict__ pos, uint8_t* __restrict__ valid,
                                                          int32_t* __restrict__ uv, int* __restrict__ status) {
    const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (f >= n) return;
    const uint8_t* e = blob + (size_t)f * wire_elem_bytes(cols);
    const uint32_t len = load_u32_unaligned(e + 13);
    if (len != (uint32_t)cols) { if (lane == 0) atomicExch(status, 1); return; }
    for (int j = lane; j < cols; j += 32) {
        const float val = __uint_as_float(load_u32_unaligned(e + 17 + 4 * j));
        desc[(size_t)f * cols + j] = narrow_x86(val);
    }
    if (lane < 3) pos[3 * (size_t)f + lane] = load_f64_unaligned(e + 17 + 4 * cols + 8 * lane);
    if (lane == 3) valid[f] = e[8] ? 1 : 0;
    if (uv && lane == 4) { uv[2 * f] = (int32_t)load_u32_unaligned(e); uv[2 * f + 1] = (int32_t)load_u32_unaligned(e + 4); }
}

// Many blobs in one launch (resume: GraphSlamNode::load re-adds every stored node, graph_slam_node.cpp:875-888).
// blockIdx.y = blob; status[0] is set to 1 if any element's descriptor length differs from its blob's.
struct WireJob { const uint8_t* blob; uint8_t* desc; double* pos; uint8_t* valid; int32_t n, cols; };
__global__ void __launch_bounds__(256) wire_decode_bulk_kernel(const WireJob* __restrict__ jobs, int* __restrict__ status) {
    const WireJob j = jobs[blockIdx.y];
    const int lane = threadIdx.x & 31;
    for (int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; f < j.n; f += (gridDim.x * blockDim.x) >> 5) {
        const uint8_t* e = j.blob + (size_t)f * wire_elem_bytes(j.cols);
        const uint32_t len = load_u32_unaligned(e + 13);
        if (len != (uint32_t)j.cols) { if (lane == 0) atomicExch(status, 1); continue; }
        for (int k = lane; k < j.cols; k += 32) j.desc[(size_t)f * j.cols + k] = narrow_x86(__uint_as_float(load_u32_unaligned(e + 17 + 4 * k)));
        if (lane < 3) j.pos[3 * (size_t)f + lane] = load_f64_unaligned(e + 17 + 4 * j.cols + 8 * lane);
        if (lane == 3) j.valid[f] = e[8] ? 1 : 0;
    }
}

// FeatureData::toMsg (sensor_data.cpp:77-122): the inverse of wire_decode_kernel, one warp per feature.  blob points at the
// first element; uv (n x 2, may be null) supplies feature_positions_2d_, keypoint_strength is -1.
__device__ __forceinline__ void store_u32_unaligned(uint8_t* p, uint32_t v) {
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}
__global__ void __launch_bounds__(256) wire_encode_kernel(const uint8_t* __restrict__ desc, int n, int cols, const double* __restrict__ pos,
                                                          const uint8_t* __restrict__ valid, const int32_t* __restrict__ uv,
                                                          uint8_t* __restrict__ blob) {
    const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (f >= n) return;
    uint8_t* e = blob + (size_t)f * wire_elem_bytes(cols);
    if (lane == 0) {
        store_u32_unaligned(e, uv ? (uint32_t)uv[2 * f] : 0u);
        store_u32_unaligned(e + 4, uv ? (uint32_t)uv[2 * f + 1] : 0u);
        e[8] = valid[f] ? 1 : 0;
        store_u32_unaligned(e + 9, __float_as_uint(-1.0f));
        store_u32_unaligned(e + 13, (uint32_t)cols);
    }
    for (int j = lane; j < cols; j += 32) store_u32_unaligned(e + 17 + 4 * j, __float_as_uint((float)desc[(size_t)f * cols + j]));
    if (lane < 3) {
        const unsigned long long b = (unsigned long long)__double_as_longlong(pos[3 * (size_t)f + lane]);
        store_u32_unaligned(e + 17 + 4 * cols + 8 * lane, (uint32_t)b);
        store_u32_unaligned(e + 17 + 4 * cols + 8 * lane + 4, (uint32_t)(b >> 32));
    }
}

}  // namespace uz
