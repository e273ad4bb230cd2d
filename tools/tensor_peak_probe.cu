/*
 * tools/tensor_peak_probe.cu -- MEASURED dense tcgen05 throughput of this GPU for the operand kinds the solver uses
 * (kind::i8 for the batched loop, kind::tf32 for the setup GEMMs) and kind::f16 (bf16) for comparison with
 * MEASURED_PEAKS.json's cuBLAS figure.  One CTA per SM, operands resident in shared memory, M = 128 per CTA
 * (cta_group::1, what the product kernels issue), all 148 SMs, wall-clock (CUDA events) over >= 0.25 s per shape so
 * the clocks are the sustained ones.  Prints one JSON object; `python tools/tensor_peaks.py` (GPU box) stores it as
 * profiles/tensor_peaks_r2.json, which bench.py uses as roofline.peak for the batched and setup legs.
 *   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I pqp-for-mpc_b200/csrc -o tools/tensor_peak_probe.bin tools/tensor_peak_probe.cu
 */
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "pqp_umma.cuh"

__device__ __forceinline__ void mma_kind(int kind, uint32_t d, uint64_t a, uint64_t b, uint32_t idesc)
{
	if (kind == 0)
		asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
	else if (kind == 2)
		asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
	else
		asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
__device__ __forceinline__ bool elect_one()
{
	uint32_t pred;
	asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
	return pred != 0;
}

/* nmma MMAs of N columns each per group, spread over `nacc` accumulators; 16 groups per commit */
__global__ void __launch_bounds__(128, 1) peak_kernel(int kind, int n, int nacc, int rounds)
{
	extern __shared__ __align__(1024) unsigned char smem[];
	__shared__ uint64_t bar;
	__shared__ uint32_t slot;
	const int tid = threadIdx.x, warp = tid / 32;
	for (int i = tid; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0u;
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
		const uint32_t id = fmt | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); /* both operands K-major */
		const uint64_t da = umma::smem_desc(base, 2048, 128), db = umma::smem_desc(base + 32 * 1024, 4096, 128);
		uint32_t phase = 0;
		for (int r = 0; r < rounds; r++) {
			if (elect_one()) {
#pragma unroll
				for (int s = 0; s < 16; s++) mma_kind(kind, tmem + (uint32_t)(s % nacc) * (uint32_t)n, da + (uint64_t)(s & 3) * 256u, db, id);
				umma::mma_commit(&bar);
			}
			__syncwarp();
			umma::mbar_wait(&bar, phase);
			phase ^= 1u;
		}
	}
	umma::tc_fence_before();
	__syncthreads();
	if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

int main()
{
	int sms = 0, clk = 0;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
	cudaFuncSetAttribute(peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	const char *names[3] = { "tf32", "i8", "bf16" };
	const int kdepth[3] = { 8, 32, 16 };
	printf("{\"sms\": %d, \"sm_clock_khz\": %d, \"how\": \"tools/tensor_peak_probe.cu: tcgen05.mma cta_group::1 M=128, operands resident in shared memory, all SMs, CUDA events over >= 0.25 s per shape\"", sms, clk);
	for (int kind = 0; kind < 3; kind++) {
		double best = 0.0;
		int best_n = 0, best_acc = 0;
		for (int n : { 64, 128, 192, 256 })
			for (int nacc : { 1, 2 }) {
				if (n * nacc > 512) continue;
				int rounds = 2000;
				float ms = 0.0f;
				for (int rep = 0; rep < 4; rep++) { /* grow until the launch lasts >= 0.25 s */
					cudaEventRecord(e0);
					peak_kernel<<<sms, 128, 96 * 1024>>>(kind, n, nacc, rounds);
					cudaEventRecord(e1);
					if (cudaEventSynchronize(e1) != cudaSuccess) {
						fprintf(stderr, "probe failed: %s\n", cudaGetErrorString(cudaGetLastError()));
						return 1;
					}
					cudaEventElapsedTime(&ms, e0, e1);
					if (ms >= 250.0f) break;
					rounds = (int)(rounds * (300.0f / (ms > 1.0f ? ms : 1.0f))) + 1;
				}
				const double ops = 2.0 * 128.0 * n * kdepth[kind] * 16.0 * rounds * sms;
				const double tops = ops / (ms * 1e-3) / 1e12;
				fprintf(stderr, "%s N=%d acc=%d: %.1f T(FL)OP/s (%.0f ms)\n", names[kind], n, nacc, tops, ms);
				if (tops > best) { best = tops; best_n = n; best_acc = nacc; }
			}
		printf(", \"%s_tops\": %.1f, \"%s_shape\": \"M=128 N=%d K=%d, %d accumulator(s)\"", names[kind], best, names[kind], best_n, kdepth[kind], best_acc);
	}
	printf("}\n");
	return 0;
}
