// a52_imdct_ab.inl - the tensor-core IMDCT experiment (round 2, north star: "tensor cores are used only if a batched
// DCT-IV-as-GEMM variant beats the FFT path on ncu-measured throughput").
//
// Two stand-alone kernels compute the SAME thing - for every 256-coefficient plane the 256 values (U[128], V[128])
// that imdct512_warp / imdct512_pair leave for the window + overlap-add stage (imdct.c:258-345) - from global memory
// to global memory, so that they can be timed and profiled against each other:
//   a52_ab_fft_kernel : the production transform (imdct512_pair: two planes packed in f32x2, radix-4 FFT in shared
//                       memory), one warp per pair of planes
//   a52_ab_mma_kernel : Y[256 x N] = M[256 x 256] . X[256 x N] on the tensor cores (mma.sync m16n8k8 TF32, three
//                       products per term - "3xTF32": M = Mhi + Mlo, X = Xhi + Xlo, Mlo.Xhi + Mhi.Xlo + Mhi.Xhi - which
//                       is what the 1e-5 PCM bar needs; plain TF32 is 1e-3), 64 planes per CTA, M's fragments
//                       streamed from L2 in fragment order, X split once into shared memory
// Included at the end of a52_decode.cu (same translation unit: the tables and imdct512_pair are there).
// Driver: tools/dev_imdct_ab.py; result and decision: profiles/r02_imdct_tc_ab.json, DESIGN.md section 4.

namespace a52 {

__global__ void __launch_bounds__(256, 2)
a52_ab_fft_kernel(const float* __restrict__ x, float* __restrict__ y, int npairs)
{
    extern __shared__ __align__(128) uint8_t smem[];
    Tables& T = *reinterpret_cast<Tables*>(smem);
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&g_tables);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&T);
        for (int i = threadIdx.x; i < (int)(sizeof(Tables) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    float* buf = reinterpret_cast<float*>(smem + align128((int)sizeof(Tables)) + warp * 2048);
    const uint32_t tab_base = smem_u32(&T), x_sa = smem_u32(buf);
    for (int p = blockIdx.x * nw + warp; p < npairs; p += gridDim.x * nw) {
        const float* xl = x + (size_t)(2 * p) * 256, *xr = xl + 256;
        // (L, R) pairs in mix2_pairs' order: pair k at ((k >> 6) & 1) * 1024 + (k >> 7) * 512 + (k & 63) * 8 bytes
        for (int k = lane; k < 256; k += 32) {
            const uint32_t o = ((k >> 6) & 1) * 1024u + (k >> 7) * 512u + (k & 63) * 8u;
            *reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(buf) + o) = make_float2(xl[k], xr[k]);
        }
        __syncwarp();
        imdct512_pair(tab_base, x_sa, lane);
        // [i] = (U_L[2i], U_R[2i], U_L[2i+1], U_R[2i+1]), [64 + i] = the same of V
        float* yl = y + (size_t)(2 * p) * 256, *yr = yl + 256;
        for (int i = lane; i < 128; i += 32) {
            const float4 v = reinterpret_cast<const float4*>(buf)[i];
            *reinterpret_cast<float2*>(yl + 2 * i) = make_float2(v.x, v.z);
            *reinterpret_cast<float2*>(yr + 2 * i) = make_float2(v.y, v.w);
        }
        __syncwarp();
    }
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int kAbPlanes = 64;            // planes per CTA (N tile)
constexpr int kAbStride = kAbPlanes + 8; // shared-memory row stride (words): conflict-free fragment loads

// afrag: [16 m-tiles][32 k-steps][32 lanes][8] = (hi a0..a3, lo a0..a3) of M in mma fragment order
__global__ void __launch_bounds__(256, 1)
a52_ab_mma_kernel(const float* __restrict__ x, float* __restrict__ y, int nplanes, const uint4* __restrict__ afrag)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t* xhi = reinterpret_cast<uint32_t*>(smem);
    uint32_t* xlo = xhi + 256 * kAbStride;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    for (int n0 = blockIdx.x * kAbPlanes; n0 < nplanes; n0 += gridDim.x * kAbPlanes) {
        __syncthreads();
        // X[k][n] = x[n0 + n][k], split into TF32 hi + lo
        for (int i = threadIdx.x; i < 256 * kAbPlanes; i += blockDim.x) {
            const int n = i >> 8, k = i & 255;
            const float v = (n0 + n < nplanes) ? x[(size_t)(n0 + n) * 256 + k] : 0.f;
            uint32_t h, l;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(v - __uint_as_float(h)));
            xhi[k * kAbStride + n] = h;
            xlo[k * kAbStride + n] = l;
        }
        __syncthreads();
        float acc[2][8][4];
#pragma unroll
        for (int m = 0; m < 2; m++)
#pragma unroll
            for (int j = 0; j < 8; j++)
#pragma unroll
                for (int r = 0; r < 4; r++) acc[m][j][r] = 0.f;
#pragma unroll 1
        for (int ks = 0; ks < 32; ks++) {
            uint32_t ah[2][4], al[2][4];
#pragma unroll
            for (int m = 0; m < 2; m++) {
                const uint4* ap = afrag + ((size_t)((2 * warp + m) * 32 + ks) * 32 + lane) * 2;
                const uint4 h = __ldg(ap), l = __ldg(ap + 1);
                ah[m][0] = h.x; ah[m][1] = h.y; ah[m][2] = h.z; ah[m][3] = h.w;
                al[m][0] = l.x; al[m][1] = l.y; al[m][2] = l.z; al[m][3] = l.w;
            }
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int o0 = (ks * 8 + t) * kAbStride + j * 8 + g, o1 = o0 + 4 * kAbStride;
                const uint32_t bh0 = xhi[o0], bh1 = xhi[o1], bl0 = xlo[o0], bl1 = xlo[o1];
#pragma unroll
                for (int m = 0; m < 2; m++) {
                    mma_tf32(acc[m][j], al[m], bh0, bh1);          // small terms first
                    mma_tf32(acc[m][j], ah[m], bl0, bl1);
                    mma_tf32(acc[m][j], ah[m], bh0, bh1);
                }
            }
        }
        // D[row][n]: c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1)
#pragma unroll
        for (int m = 0; m < 2; m++)
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int row = (2 * warp + m) * 16 + g, n = n0 + j * 8 + 2 * t;
                if (n < nplanes) { y[(size_t)n * 256 + row] = acc[m][j][0]; y[(size_t)n * 256 + row + 8] = acc[m][j][2]; }
                if (n + 1 < nplanes) { y[(size_t)(n + 1) * 256 + row] = acc[m][j][1]; y[(size_t)(n + 1) * 256 + row + 8] = acc[m][j][3]; }
            }
    }
}

// FP32 peak of the device, measured (SURVEY.md section 8d asks for it instead of the nominal figure): eight independent
// FMA chains per thread, nothing else in the loop.
__global__ void __launch_bounds__(256) a52_ab_fma_kernel(float* out, int iters, float a, float b)
{
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = (float)(threadIdx.x + k);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = fmaf(v[k], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) s += v[k];
    if (s == 12345.678f) out[0] = s;             // keeps the chains alive
}

}  // namespace a52

#pragma GCC visibility push(default)
extern "C" {

// variant 0: FFT (production transform), 1: tensor cores (3xTF32 mma.sync).  x, y: device float [nplanes][256]
// (nplanes even); afrag: device, 16*32*32*2 uint4 (variant 1).  Launches on `cuda_stream`; returns 0 or -1.
int a52_ab_imdct(a52_batch_t* ctx, int variant, const float* x, float* y, int nplanes, const void* afrag, void* cuda_stream)
{
    using namespace a52;
    if (!ctx || nplanes <= 0 || (nplanes & 1)) return -1;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (variant == 0) {
        const size_t smem = (size_t)align128((int)sizeof(Tables)) + 8 * 2048;
        cudaFuncSetAttribute(a52_ab_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        a52_ab_fft_kernel<<<ctx->num_sms * 2, 256, smem, st>>>(x, y, nplanes / 2);
    } else {
        const size_t smem = (size_t)2 * 256 * kAbStride * 4;
        cudaFuncSetAttribute(a52_ab_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        a52_ab_mma_kernel<<<ctx->num_sms, 256, smem, st>>>(x, y, nplanes, (const uint4*)afrag);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// Measured FP32 FMA throughput of the device in TFLOP/s (2 flops per FMA), CUDA events around 10 launches.
double a52_ab_fp32_peak(a52_batch_t* ctx)
{
    using namespace a52;
    if (!ctx) return -1.0;
    float* d = nullptr;
    if (cudaMalloc(&d, 256) != cudaSuccess) return -1.0;
    const int iters = 20000, blocks = ctx->num_sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    a52_ab_fma_kernel<<<blocks, 256>>>(d, iters, 1.0000001f, 1e-9f);
    cudaEventRecord(e0);
    for (int r = 0; r < 10; r++) a52_ab_fma_kernel<<<blocks, 256>>>(d, iters, 1.0000001f, 1e-9f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    if (ms <= 0) return -1.0;
    return 10.0 * (double)blocks * 256 * 8 * (double)iters * 2.0 / (ms * 1e-3) / 1e12;
}

}  // extern "C"
#pragma GCC visibility pop
