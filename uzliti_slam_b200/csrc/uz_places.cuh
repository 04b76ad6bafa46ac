// uz_places.cuh — candidate generation on the device: LSH-bucket voting over the resident
// keyframe store, replacing LshSetRecognizer / FastLshSet
// (/root/reference/place_recognition/src/lsh_set_recognizer.cpp:46-94 searchAndAddPlaceImpl, :121-165 searchImpl,
//  :188-305 FastLshTable / FastLshSet) and the neighbour filter of PlaceRecognizer::searchAndAddPlace
// (/root/reference/place_recognition/src/place_recognizer.cpp:91-104: live place, |dt| > 5 s, first k).
//
// The reference keeps 8 hash tables (descriptor bytes [4k, 4k+4) -> list of place indices, one entry per descriptor
// row) and, per new keyframe, walks 8 buckets per row, incrementing one counter per listed place.  Here:
//   * ONE open-addressing table in HBM keyed by (k, 32-bit key); a slot heads a linked list of nodes (place, next)
//     living in an append-only node array.  Insertion is one atomicCAS probe + one atomicExch; nothing is ever moved,
//     so the structure is incremental like the reference's (searchAndAddPlace keyframe by keyframe) and batches of
//     keyframes insert in one launch.
//   * sequential semantics without sequential execution: the i-th keyframe of a batch must only see places < i.
//     Every node carries its place index, all keyframes of the batch are inserted first, and a query counts a node
//     only if node.place < the query's own place index.
//   * votes land in a dense row per query camera (uint32 [n_places], L2-resident atomics); one CTA per row then
//     filters (votes >= 8T, live, |dt| > 5 s) and sorts the survivors by (votes desc, place asc) in shared memory.
// Work is random 8..32-byte accesses into HBM/L2: the bound is memory transactions, not arithmetic.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace uz {

struct PlaceSlot { unsigned long long tag; uint32_t head; uint32_t pad; };      // tag 0 = empty; head = node index + 1
struct PlaceNode { uint32_t place; uint32_t next; };                            // next = node index + 1, 0 = end

// one FEATURE sensor of one keyframe taking part in a launch
struct PlaceCam {
    const uint32_t* raw;      // n rows of 8 << wide_shift words, descriptor bytes as given (store "raw" layout); the
                              // tables key on bytes [0, 32) of a row whatever its width (lsh_set_recognizer.cpp:243)
    int32_t n;
    int32_t place;            // place index of the owning keyframe
    uint32_t node_base;       // first node of this camera in the node array (insert launches)
    int32_t insert_filtered;  // 1: matchAndAdd's popcount filter applies to the inserted keys; 0: FastLshSet::add
    int32_t query_filtered;   // 1: matchAndAdd (popcount filter on the query keys); 0: FastLshSet::match
    int32_t place_limit;      // count only nodes with place < place_limit
    int32_t row;              // votes row of this camera in the launch
    int32_t wide_shift;       // 0: 32-byte rows, 1: 64-byte rows
    long long stamp_ns;       // time stamp of the querying keyframe (pr_time_map_[id])
};

__host__ __device__ __forceinline__ unsigned long long place_tag(int k, uint32_t key) {
    return ((unsigned long long)(k + 1) << 32) | key;
}
__device__ __forceinline__ uint32_t place_hash(unsigned long long tag) {
    unsigned long long h = tag * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29;
    h *= 0xBF58476D1CE4E5B9ull;
    return (uint32_t)(h >> 32);
}

// rows[i] = (camera, row) enumerated on the fly: blockIdx.y = camera, threads cover row*8 + k
__global__ void __launch_bounds__(256) places_insert_kernel(const PlaceCam* __restrict__ cams, PlaceSlot* __restrict__ slots,
                                                            uint32_t slot_mask, PlaceNode* __restrict__ nodes, int min_key_bits) {
    const PlaceCam cam = cams[blockIdx.y];
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < cam.n * 8; e += gridDim.x * blockDim.x) {
        const int k = e & 7;
        const uint32_t key = cam.raw[((e >> 3) << (3 + cam.wide_shift)) + k];     // word k of row e/8 == bytes [4k, 4k+4) little endian
        const uint32_t node = cam.node_base + (uint32_t)e;
        if (cam.insert_filtered && __popc(key) <= min_key_bits) { nodes[node] = PlaceNode{0xFFFFFFFFu, 0u}; continue; }
        const unsigned long long tag = place_tag(k, key);
        uint32_t s = place_hash(tag) & slot_mask;
        while (true) {
            const unsigned long long cur = atomicCAS(&slots[s].tag, 0ull, tag);
            if (cur == 0ull || cur == tag) break;
            s = (s + 1) & slot_mask;
        }
        const uint32_t prev = atomicExch(&slots[s].head, node + 1u);
        nodes[node] = PlaceNode{(uint32_t)cam.place, prev};
    }
}

__global__ void __launch_bounds__(256) places_vote_kernel(const PlaceCam* __restrict__ cams, const PlaceSlot* __restrict__ slots,
                                                          uint32_t slot_mask, const PlaceNode* __restrict__ nodes,
                                                          uint32_t* __restrict__ votes, int n_places, int min_key_bits) {
    const PlaceCam cam = cams[blockIdx.y];
    uint32_t* __restrict__ row = votes + (size_t)cam.row * n_places;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < cam.n * 8; e += gridDim.x * blockDim.x) {
        const int k = e & 7;
        const uint32_t key = cam.raw[((e >> 3) << (3 + cam.wide_shift)) + k];
        if (cam.query_filtered && __popc(key) <= min_key_bits) continue;
        const unsigned long long tag = place_tag(k, key);
        uint32_t s = place_hash(tag) & slot_mask;
        uint32_t head = 0;
        while (true) {
            const unsigned long long cur = slots[s].tag;
            if (cur == tag) { head = slots[s].head; break; }
            if (cur == 0ull) break;
            s = (s + 1) & slot_mask;
        }
        while (head) {
            const PlaceNode nd = nodes[head - 1];
            // a node linked by the running insert of a LATER place may not have its payload yet when inserts and
            // votes of different launches overlap; launches are stream ordered here, so payloads are always there
            if ((int)nd.place < cam.place_limit) atomicAdd(&row[nd.place], 1u);
            head = nd.next;
        }
    }
}

struct PlaceSelectParams {
    uint32_t min_votes;           // max(1, ceil(8 T)): all_matches[i] > 0 && votes/8 >= T
    long long min_gap_ns;         // 5 s (place_recognizer.cpp:94)
    int32_t k;                    // k_nearest_neighbors
    int32_t n_places;
};

constexpr int kSelectCap = 4096;  // candidates sorted on chip; beyond that the exact k-pass fallback runs

// One CTA per votes row.  out[row*k + r] = r-th best place (or -1), out_votes likewise.
__global__ void __launch_bounds__(256) places_select_kernel(const PlaceCam* __restrict__ cams, const uint32_t* __restrict__ votes,
                                                            const long long* __restrict__ stamp_ns, const uint8_t* __restrict__ live,
                                                            PlaceSelectParams prm, int32_t* __restrict__ out, uint32_t* __restrict__ out_votes) {
    __shared__ unsigned long long cand[kSelectCap];
    __shared__ int s_count;
    __shared__ unsigned long long s_best[8];
    const PlaceCam cam = cams[blockIdx.x];
    const uint32_t* __restrict__ row = votes + (size_t)cam.row * prm.n_places;
    const long long t_me = cam.stamp_ns;
    const int limit = min(cam.place_limit, prm.n_places);
    const int tid = threadIdx.x;
    int32_t* o = out + (size_t)cam.row * prm.k;
    uint32_t* ov = out_votes + (size_t)cam.row * prm.k;
    if (tid == 0) s_count = 0;
    __syncthreads();
    auto key_of = [&](int j) -> unsigned long long {
        const uint32_t v = row[j];
        if (v < prm.min_votes || !live[j]) return 0ull;
        long long dt = stamp_ns[j] - t_me;
        if (dt < 0) dt = -dt;
        if (dt <= prm.min_gap_ns) return 0ull;
        return ((unsigned long long)v << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)j);    // bigger = better
    };
    for (int j = tid; j < limit; j += blockDim.x) {
        const unsigned long long key = key_of(j);
        if (key) {
            const int p = atomicAdd(&s_count, 1);
            if (p < kSelectCap) cand[p] = key;
        }
    }
    __syncthreads();
    const int count = s_count;
    if (count <= kSelectCap) {
        int n2 = 1;
        while (n2 < count) n2 <<= 1;
        for (int i = count + tid; i < n2; i += blockDim.x) cand[i] = 0ull;
        __syncthreads();
        for (int size = 2; size <= n2; size <<= 1)                     // bitonic sort, descending
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int i = tid; i < n2; i += blockDim.x) {
                    const int j = i ^ stride;
                    if (j > i) {
                        const bool desc = (i & size) == 0;
                        const unsigned long long a = cand[i], b = cand[j];
                        if ((a < b) == desc) { cand[i] = b; cand[j] = a; }
                    }
                }
                __syncthreads();
            }
        for (int r = tid; r < prm.k; r += blockDim.x) {
            const bool has = r < count;
            o[r] = has ? (int32_t)(0xFFFFFFFFu - (uint32_t)(cand[r] & 0xFFFFFFFFull)) : -1;
            ov[r] = has ? (uint32_t)(cand[r] >> 32) : 0u;
        }
    } else {
        // exact fallback for pathological rows (thousands of places over threshold): k passes of a block-wide maximum
        // over the keys strictly below the previous winner
        unsigned long long bound = ~0ull;
        for (int r = 0; r < prm.k; ++r) {
            unsigned long long best = 0ull;
            for (int j = tid; j < limit; j += blockDim.x) {
                const unsigned long long key = key_of(j);
                if (key < bound && key > best) best = key;
            }
            for (int off = 16; off > 0; off >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, off);
                if (other > best) best = other;
            }
            __syncthreads();
            if ((tid & 31) == 0) s_best[tid >> 5] = best;
            __syncthreads();
            best = s_best[0];
            for (int w = 1; w < 8; ++w) if (s_best[w] > best) best = s_best[w];
            if (tid == 0) {
                o[r] = best ? (int32_t)(0xFFFFFFFFu - (uint32_t)(best & 0xFFFFFFFFull)) : -1;
                ov[r] = (uint32_t)(best >> 32);
            }
            bound = best ? best : 0ull;          // 0: nothing left, the remaining ranks stay empty
        }
    }
}

// growth: re-link every live node into a bigger slot array (tags are recomputed from the descriptors, so nodes stay 8 B)
__global__ void __launch_bounds__(256) places_relink_kernel(const PlaceCam* __restrict__ cams, PlaceSlot* __restrict__ slots,
                                                            uint32_t slot_mask, PlaceNode* __restrict__ nodes) {
    const PlaceCam cam = cams[blockIdx.y];
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < cam.n * 8; e += gridDim.x * blockDim.x) {
        const uint32_t node = cam.node_base + (uint32_t)e;
        if (nodes[node].place == 0xFFFFFFFFu) continue;                  // filtered out at insertion
        const unsigned long long tag = place_tag(e & 7, cam.raw[((e >> 3) << (3 + cam.wide_shift)) + (e & 7)]);
        uint32_t s = place_hash(tag) & slot_mask;
        while (true) {
            const unsigned long long cur = atomicCAS(&slots[s].tag, 0ull, tag);
            if (cur == 0ull || cur == tag) break;
            s = (s + 1) & slot_mask;
        }
        nodes[node].next = atomicExch(&slots[s].head, node + 1u);
    }
}

}  // namespace uz
