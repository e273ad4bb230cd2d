/*
 * tools/mma_ts_probe.cu -- pace of tcgen05.mma.kind::i8 with the A operand in TMEM (fed by tcgen05.cp from shared memory)
 * against the SS form, at the shapes of the batched kernel's K step (N = 96, 64, 32).  Timing only (operand values arbitrary).
 *   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I pqp-for-mpc_b200/csrc -o tools/mma_ts_probe.bin tools/mma_ts_probe.cu
 */
#include <cstdio>
#include <cuda_runtime.h>
#include "pqp_umma.cuh"

__device__ __forceinline__ bool elect_one()
{
	uint32_t pred;
	asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
	return pred != 0;
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc)
{
	asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc)
{
	asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc) : "memory");
}
__device__ __forceinline__ void cp_128x256b(uint32_t taddr, uint64_t sdesc)
{
	asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}

/* mode 0: SS (A from smem); 1: TS, A tiles already in TMEM (no copies); 2: TS with one tcgen05.cp per A tile */
__global__ void __launch_bounds__(128, 1) probe(int mode, int n0, int n1, int n2, int rounds, long long *out)
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
		const uint32_t fmt = (2u << 4) | (0u << 7) | (1u << 10) | (1u << 16);
		const uint32_t i0 = fmt | ((uint32_t)(n0 >> 3) << 17) | ((128u >> 4) << 24);
		const uint32_t i1 = fmt | ((uint32_t)(n1 >> 3) << 17) | ((128u >> 4) << 24);
		const uint32_t i2 = fmt | ((uint32_t)(n2 >> 3) << 17) | ((128u >> 4) << 24);
		const uint64_t db = umma::smem_desc(base + 64 * 1024, 128, 1024);
		const uint64_t da = umma::smem_desc(base, 2048, 128);
		const uint32_t a_t = tmem + 256; /* A tiles: 8 columns each, a ring of 24 */
		uint32_t phase = 0;
		long long t0 = clock64();
		for (int r = 0; r < rounds; r++) {
			if (elect_one()) {
#pragma unroll
				for (int s = 0; s < 16; s++) {
					const uint64_t a0 = da + (uint64_t)(s % 5) * 768u;
					const uint32_t at = a_t + (uint32_t)(s % 8) * 24u;
					if (mode == 0) {
						mma_ss(tmem, a0, db, i0);
						mma_ss(tmem + n0 - n1, a0 + 256, db, i1);
						mma_ss(tmem + n0 - n2, a0 + 512, db, i2);
					} else {
						if (mode == 2) {
							cp_128x256b(at, a0);
							cp_128x256b(at + 8, a0 + 256);
							cp_128x256b(at + 16, a0 + 512);
						}
						mma_ts(tmem, at, db, i0);
						mma_ts(tmem + n0 - n1, at + 8, db, i1);
						mma_ts(tmem + n0 - n2, at + 16, db, i2);
					}
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
	const char *names[] = { "SS (A from shared memory)", "TS, A resident in TMEM (no copies)", "TS + tcgen05.cp 128x256b per A tile" };
	int shapes[][3] = { { 96, 64, 32 }, { 192, 128, 64 }, { 32, 32, 32 }, { 96, 96, 96 } };
	for (auto &sh : shapes)
		for (int mode = 0; mode < 3; mode++) {
			const int rounds = 200;
			probe<<<148, 128, 160 * 1024>>>(mode, sh[0], sh[1], sh[2], rounds, d);
			cudaError_t e = cudaDeviceSynchronize();
			if (e != cudaSuccess) {
				printf("%s: %s\n", names[mode], cudaGetErrorString(e));
				return 1;
			}
			long long h[148];
			cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
			long long mx = 0;
			for (int i = 0; i < 148; i++) mx = h[i] > mx ? h[i] : mx;
			printf("N = %3d+%3d+%3d  %-40s %8.1f cycles per K step\n", sh[0], sh[1], sh[2], names[mode], (double)mx / (rounds * 16.0));
		}
	return 0;
}
