/*
 * tools/mma_rate_probe.cu -- issue rate of tcgen05.mma (kind::tf32 vs kind::i8) at the shapes the batched kernels use.
 *   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I pqp-for-mpc_b200/csrc -o /tmp/mma_rate_probe tools/mma_rate_probe.cu
 *   /tmp/mma_rate_probe
 * One CTA per SM, operands resident in shared memory (no streaming): cycles per MMA = the tensor pipe's own pace
 * (including its shared-memory operand fetch).  Experiment tool, not part of the product.
 */
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "pqp_umma.cuh"

__device__ __forceinline__ void mma_any(int kind, uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
	if (kind == 0)
		asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a),
			     "l"(b), "r"(idesc), "r"(acc)
			     : "memory");
	else if (kind == 2)
		asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a),
			     "l"(b), "r"(idesc), "r"(acc)
			     : "memory");
	else
		asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a),
			     "l"(b), "r"(idesc), "r"(acc)
			     : "memory");
}

__device__ __forceinline__ bool elect_one()
{
	uint32_t pred;
	asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
	return pred != 0;
}

/* warp-uniform issue loop, descriptors precomputed: the pace measured is the tensor pipe's, not the issuing thread's */
__global__ void __launch_bounds__(128, 1) probe(int kind, int n0, int n1, int n2, int bmajor_mn, int rounds, int distinct_a, int swz, int alt_d,
						long long *out)
{
	extern __shared__ __align__(1024) unsigned char smem[];
	__shared__ uint64_t bar;
	__shared__ uint32_t slot;
	const int tid = threadIdx.x, warp = tid / 32;
	for (int i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u;
	if (tid == 0) {
		umma::mbar_init(&bar, 1);
		umma::mbar_fence_init();
	}
	if (warp == 0) umma::tmem_alloc(&slot, 512);
	umma::fence_proxy_async();
	umma::tc_fence_before();
	__syncthreads();
	umma::tc_fence_after();
	const uint32_t tmem = slot;
	if (warp == 1) {
		const uint32_t base = umma::smem_addr(smem);
		const uint32_t fmt = kind == 0 ? ((1u << 4) | (2u << 7) | (2u << 10)) : (kind == 2 ? ((1u << 4) | (1u << 7) | (1u << 10)) : ((2u << 4) | (0u << 7) | (1u << 10)));
		const uint32_t i0 = fmt | ((uint32_t)bmajor_mn << 16) | ((uint32_t)(n0 >> 3) << 17) | ((128u >> 4) << 24);
		const uint32_t i1 = fmt | ((uint32_t)bmajor_mn << 16) | ((uint32_t)(n1 >> 3) << 17) | ((128u >> 4) << 24);
		const uint32_t i2 = fmt | ((uint32_t)bmajor_mn << 16) | ((uint32_t)(n2 >> 3) << 17) | ((128u >> 4) << 24);
		const uint32_t b_addr = base + 64 * 1024;
		const uint64_t lt = (uint64_t)swz << 61; /* layout type: 0 none, 2 SWIZZLE_128B */
		const uint64_t db = (bmajor_mn ? umma::smem_desc(b_addr, 128, 1024) : (swz ? umma::smem_desc(b_addr, 16, 1024) : umma::smem_desc(b_addr, 4096, 128))) | lt;
		const uint64_t da = (swz ? umma::smem_desc(base, 16, 1024) : umma::smem_desc(base, 2048, 128)) | lt;
		const uint64_t step = distinct_a ? (12288u >> 4) : 0u;
		uint32_t phase = 0;
		long long t0 = clock64();
		for (int r = 0; r < rounds; r++) {
			if (elect_one()) {
#pragma unroll
				for (int s = 0; s < 16; s++) {
					const uint32_t dd = tmem + (alt_d ? (uint32_t)(s & 1) * 256u : 0u);
					const uint64_t a0 = da + (uint64_t)(s % 5) * step;
					mma_any(kind, dd, a0, db, i0, 1u);
					if (n1) mma_any(kind, dd + n0 - n1, a0 + 256, db, i1, 1u);
					if (n2) mma_any(kind, dd + n0 - n2, a0 + 512, db, i2, 1u);
				}
				umma::mma_commit(&bar);
			}
			__syncwarp();
			umma::mbar_wait(&bar, phase);
			phase ^= 1u;
		}
		long long t1 = clock64();
		if (tid == 32) out[blockIdx.x] = t1 - t0;
	}
	umma::tc_fence_before();
	__syncthreads();
	if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

int main()
{
	long long *d;
	cudaMalloc(&d, 148 * sizeof(long long));
	cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
	struct { const char *name; int kind, n0, n1, n2, mn; } cases[] = {
		{ "bf16 N=64 (K=16)", 2, 64, 0, 0, 0 },
		{ "bf16 N=128", 2, 128, 0, 0, 0 },
		{ "bf16 N=256", 2, 256, 0, 0, 0 },
		{ "tf32 N=32 (K=8)", 0, 32, 0, 0, 0 },
		{ "tf32 N=64 (K=8)", 0, 64, 0, 0, 0 },
		{ "tf32 N=128", 0, 128, 0, 0, 0 },
		{ "tf32 N=256", 0, 256, 0, 0, 0 },
		{ "tf32 N=64+32 (3xTF32 step)", 0, 64, 32, 0, 0 },
		{ "i8 N=32 (K=32) B K-major", 1, 32, 0, 0, 0 },
		{ "i8 N=64 B K-major", 1, 64, 0, 0, 0 },
		{ "i8 N=128 B K-major", 1, 128, 0, 0, 0 },
		{ "i8 N=256 B K-major", 1, 256, 0, 0, 0 },
		{ "i8 N=32 B MN-major", 1, 32, 0, 0, 1 },
		{ "i8 N=64 B MN-major", 1, 64, 0, 0, 1 },
		{ "i8 N=96 B MN-major", 1, 96, 0, 0, 1 },
		{ "i8 N=128 B MN-major", 1, 128, 0, 0, 1 },
		{ "i8 N=192 B MN-major", 1, 192, 0, 0, 1 },
		{ "i8 N=256 B MN-major", 1, 256, 0, 0, 1 },
		{ "i8 96+64+32 (NB=32 step) MN", 1, 96, 64, 32, 1 },
		{ "i8 96+64 MN", 1, 96, 64, 0, 1 },
		{ "i8 96+96 MN", 1, 96, 96, 0, 1 },
		{ "i8 96+96+96 MN", 1, 96, 96, 96, 1 },
		{ "i8 32+32+32 MN", 1, 32, 32, 32, 1 },
		{ "i8 120+80+40 (NB=40 step) MN", 1, 120, 80, 40, 1 },
		{ "i8 144+96+48 (NB=48 step) MN", 1, 144, 96, 48, 1 },
		{ "i8 240+160+80 (NB=80 step) MN", 1, 240, 160, 80, 1 },
		{ "i8 256+256+256 MN", 1, 256, 256, 256, 1 },
		{ "i8 192+128+64 (NB=64 step) MN", 1, 192, 128, 64, 1 },
		{ "i8 96+64+32 (NB=32 step) K-major", 1, 96, 64, 32, 0 },
	};
	for (int grid : { 148 })
		for (auto &c : cases)
			for (int variant = 0; variant < 4; variant++) {
				const int swz = (variant & 1) ? 2 : 0, alt = (variant & 2) ? 1 : 0;
				if (swz && c.mn) continue;
				if (alt && c.n0 > 256) continue;
				const int rounds = 200;
				probe<<<grid, 128, 160 * 1024>>>(c.kind, c.n0, c.n1, c.n2, c.mn, rounds, alt, swz, alt, d);
				cudaError_t e = cudaDeviceSynchronize();
				if (e != cudaSuccess) {
					printf("%s: %s\n", c.name, cudaGetErrorString(e));
					return 1;
				}
				long long h[148];
				cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
				long long mx = 0;
				for (int i = 0; i < grid; i++) mx = h[i] > mx ? h[i] : mx;
				const int per_step = 1 + (c.n1 != 0) + (c.n2 != 0);
				printf("grid %3d  %-36s %s %s: %8.1f cycles per K step (%d MMA), %7.1f per MMA\n", grid, c.name, swz ? "SW128" : "NOSWZ",
				       alt ? "5 A tiles, 2 accumulators" : "1 A tile, 1 accumulator   ", (double)mx / (rounds * 16.0), per_step, (double)mx / (rounds * 16.0 * per_step));
			}
	return 0;
}
