/*
 * tools/mma_2cta_probe.cu -- semantics of tcgen05.mma.cta_group::2.kind::i8 (a CTA pair as one MMA: M = 256, each CTA supplies
 * its own 128 rows of A and HALF of the B rows; each CTA's TMEM receives its 128 rows x all N columns).  Checks which D columns
 * the two halves of B land in.  Experiment tool for the next batched-kernel design, not part of the product.
 *   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I pqp-for-mpc_b200/csrc -o tools/mma_2cta_probe.bin tools/mma_2cta_probe.cu
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include "pqp_umma.cuh"

#define NH 32 /* B rows per CTA */

__device__ __forceinline__ uint32_t cta_rank()
{
	uint32_t r;
	asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
	return r;
}
__device__ __forceinline__ void cluster_sync()
{
	asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
	asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(128, 1) probe(const unsigned char *A, const signed char *B, int *D)
{
	__shared__ __align__(128) unsigned char a_s[128 * 32];
	__shared__ __align__(128) signed char b_s[NH * 32];
	__shared__ uint64_t bar;
	__shared__ uint32_t slot;
	const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
	const uint32_t rank = cta_rank();
	/* canonical K-major no-swizzle tiles: (k/16)*LBO + (r/8)*128 + (r%8)*16 + k%16 */
	for (int e = tid; e < 128 * 32; e += 128) {
		const int r = e / 32, k = e % 32;
		a_s[(k / 16) * 2048 + (r / 8) * 128 + (r % 8) * 16 + k % 16] = A[(rank * 128 + r) * 32 + k];
	}
	for (int e = tid; e < NH * 32; e += 128) {
		const int n = e / 32, k = e % 32;
		b_s[(k / 16) * (NH * 16) + (n / 8) * 128 + (n % 8) * 16 + k % 16] = B[(rank * NH + n) * 32 + k];
	}
	if (tid == 0) {
		umma::mbar_init(&bar, 1);
		umma::mbar_fence_init();
	}
	umma::fence_proxy_async();
	if (warp == 0 && rank == 0) { /* the leader's warp allocates the same columns in both CTAs */
		asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(umma::smem_addr(&slot)), "r"(64u) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
	}
	umma::tc_fence_before();
	__syncthreads();
	cluster_sync();
	umma::tc_fence_after();
	uint32_t tmem = slot;
	if (rank != 0) tmem = 0; /* the pair's allocation starts at the same column in both CTAs; the probe launches alone */
	if (rank == 0 && tid == 0) {
		const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)((2 * NH) >> 3) << 17) | ((256u >> 4) << 24);
		const uint64_t da = umma::smem_desc(umma::smem_addr(a_s), 2048, 128);
		const uint64_t db = umma::smem_desc(umma::smem_addr(b_s), NH * 16, 128);
		asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 1;\n\ttcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da),
			     "l"(db), "r"(idesc)
			     : "memory");
		asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
				     umma::smem_addr(&bar)),
			     "h"((uint16_t)3)
			     : "memory");
	}
	umma::mbar_wait(&bar, 0);
	umma::tc_fence_after();
	/* every CTA reads its 128 lanes x 64 columns */
	for (int c0 = 0; c0 < 2 * NH; c0 += 16) {
		float v[16];
		umma::tmem_ld16(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0, v);
		for (int j = 0; j < 16; j++) D[((rank * 128) + 32 * warp + lane) * (2 * NH) + c0 + j] = __float_as_int(v[j]);
	}
	umma::tc_fence_before();
	__syncthreads();
	cluster_sync();
	if (warp == 0 && rank == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

int main()
{
	unsigned char hA[256 * 32];
	signed char hB[2 * NH * 32];
	srand(1);
	for (auto &x : hA) x = (unsigned char)(rand() % 200);
	for (auto &x : hB) x = (signed char)(rand() % 200 - 100);
	unsigned char *dA;
	signed char *dB;
	int *dD;
	cudaMalloc(&dA, sizeof hA);
	cudaMalloc(&dB, sizeof hB);
	cudaMalloc(&dD, 256 * 2 * NH * sizeof(int));
	cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice);
	cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
	cudaMemset(dD, 0xFF, 256 * 2 * NH * sizeof(int));
	cudaLaunchConfig_t cfg;
	memset(&cfg, 0, sizeof cfg);
	cfg.gridDim = dim3(2);
	cfg.blockDim = dim3(128);
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = 2;
	attr[0].val.clusterDim.y = 1;
	attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr;
	cfg.numAttrs = 1;
	cudaError_t e = cudaLaunchKernelEx(&cfg, probe, (const unsigned char *)dA, (const signed char *)dB, dD);
	if (e == cudaSuccess) e = cudaDeviceSynchronize();
	if (e != cudaSuccess) {
		printf("launch: %s\n", cudaGetErrorString(e));
		return 1;
	}
	static int hD[256 * 2 * NH];
	cudaMemcpy(hD, dD, sizeof hD, cudaMemcpyDeviceToHost);
	/* expected: D[r][n] = sum_k A[r][k] * B[n][k], n over [CTA0's rows | CTA1's rows] */
	int bad = 0, swapped = 0;
	for (int r = 0; r < 256; r++)
		for (int n = 0; n < 2 * NH; n++) {
			long long s = 0, s2 = 0;
			for (int k = 0; k < 32; k++) {
				s += (long long)hA[r * 32 + k] * hB[n * 32 + k];
				s2 += (long long)hA[r * 32 + k] * hB[((n + NH) % (2 * NH)) * 32 + k];
			}
			if (hD[r * 2 * NH + n] != (int)s) bad++;
			if (hD[r * 2 * NH + n] == (int)s2) swapped++;
		}
	printf("cta_group::2 i8 MMA, M=256 N=%d K=32: %d of %d entries differ from [CTA0 rows | CTA1 rows] order; %d match the swapped order\n", 2 * NH, bad,
	       256 * 2 * NH, swapped);
	printf("D[0][0..3] = %d %d %d %d   D[128][0..3] = %d %d %d %d\n", hD[0], hD[1], hD[2], hD[3], hD[128 * 2 * NH], hD[128 * 2 * NH + 1],
	       hD[128 * 2 * NH + 2], hD[128 * 2 * NH + 3]);
	return bad ? 2 : 0;
}
