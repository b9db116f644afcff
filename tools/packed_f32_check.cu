// Packed-FP32 (FADD2 / FMUL2 / FFMA2) normal transform against the scalar one, bit for bit, on 16.7 M Threefry draws (needs a GPU):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false tools/packed_f32_check.cu -o /tmp/packed_f32_check && /tmp/packed_f32_check
// (how the ptxas contraction of mul.rn.f32x2 + add.rn.f32x2 was found: profiles/r2_notes.md section 1)
#include <cstdio>
#include "../qdax_b200/csrc/qdx_math.cuh"
__global__ void k(uint32_t seed, int n, int* nbad, uint32_t* bad) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t b[4];
    for (int j = 0; j < 4; ++j) b[j] = qdx_bits32(QdxKey{seed, 7u}, (uint64_t)i * 4 + j);
    float o[4];
    qdx_normal4_from_bits(b, o);
    for (int j = 0; j < 4; ++j) {
        float ref = qdx_normal_from_bits(b[j]);
        if (__float_as_uint(ref) != __float_as_uint(o[j])) {
            int s = atomicAdd(nbad, 1);
            if (s == 0) { for (int q = 0; q < 4; ++q) bad[40 + q] = b[q]; } atomicAdd((int*)&bad[48 + j], 1); if (s < 8) { bad[4*s] = b[j]; bad[4*s+1] = __float_as_uint(ref); bad[4*s+2] = __float_as_uint(o[j]); bad[4*s+3] = b[j ^ 1]; if (s == 0) bad[63] = j; }
        }
    }
    // stage-by-stage on element 0/1 of the first mismatch is done on the host side from the bits
}
__global__ void four(uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3, float* out) {
    uint32_t b[4] = {b0, b1, b2, b3}; float o[4];
    qdx_normal4_from_bits(b, o);
    if (threadIdx.x == 0) for (int q = 0; q < 4; ++q) { out[q] = o[q]; out[4 + q] = qdx_normal_from_bits(b[q]); }
}
int main() {
    int *nbad; uint32_t* bad; cudaMallocManaged(&nbad, 4); cudaMallocManaged(&bad, 64 * 4); *nbad = 0;
    const int n = 1 << 22;
    k<<<(n + 255) / 256, 256>>>(1234u, n, nbad, bad);
    cudaDeviceSynchronize();
    printf("mismatches: %d of %d\n", *nbad, 4 * n);
    for (int s = 0; s < (*nbad < 8 ? *nbad : 8); ++s) printf("  bits %08x ref %08x got %08x lane %u\n", bad[4*s], bad[4*s+1], bad[4*s+2], bad[4*s+3]);
    if (*nbad) {
        float* out; cudaMallocManaged(&out, 256);
        printf("  mismatches per element j: %u %u %u %u\n", bad[48], bad[49], bad[50], bad[51]);
        four<<<1, 32>>>(bad[40], bad[41], bad[42], bad[43], out + 24); cudaDeviceSynchronize();
        for (int q = 0; q < 4; ++q) printf("  normal4 out[%d] = %08x   single = %08x\n", q, *(uint32_t*)&out[24 + q], *(uint32_t*)&out[28 + q]);
    }
    return 0;
}
