// Probe of cp.async.bulk.tensor.2d ... tile::gather4 on sm_100a: which box shape the tensor map needs, how many bytes one
// instruction delivers and where the four rows land in (swizzled) shared memory.  Build: nvcc -arch=sm_100a -o gather4_probe gather4_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap map, int r0, int r1, int r2, int r3, uint32_t expect, double* out, int* status) {
	extern __shared__ __align__(1024) unsigned char sm[];
	uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 4096);
	double* dst = reinterpret_cast<double*>(sm);
	for (int i = threadIdx.x; i < 512; i += blockDim.x) dst[i] = -1.0;
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
		asm volatile("fence.mbarrier_init.release.cluster;");
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(expect) : "memory");
		asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(s32(dst)), "l"(&map), "r"(0),
		             "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(s32(bar))
		             : "memory");
		long long t0 = clock64();
		uint32_t ok = 0;
		while (!ok && clock64() - t0 < 20000000LL) {
			asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(s32(bar)) : "memory");
		}
		*status = ok ? 1 : -1;
	}
	__syncthreads();
	for (int i = threadIdx.x; i < 512; i += blockDim.x) out[i] = dst[i];
}

typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
	const int N = 1024, W = 16;
	std::vector<double> h((size_t)N * W);
	for (int r = 0; r < N; r++)
		for (int c = 0; c < W; c++) h[(size_t)r * W + c] = r * 100.0 + c;
	double *d, *out;
	int* st;
	cudaMalloc(&d, h.size() * 8);
	cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
	cudaMalloc(&out, 512 * 8);
	cudaMalloc(&st, 4);
	void* fn = nullptr;
	cudaDriverEntryPointQueryResult q;
	cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
	Enc enc = (Enc)fn;
	cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192);
	const int boxes[2] = {1, 4};
	const uint32_t expects[2] = {512, 2048};
	for (int wi = 0; wi < 2; wi++) {   // record width: 16 doubles (128B swizzle), 8 doubles (64B swizzle)
		const int w = wi == 0 ? 16 : 8;
		for (int b = 0; b < 2; b++) {
			CUtensorMap map;
			cuuint64_t gdim[2] = {(cuuint64_t)w, (cuuint64_t)(N * W / w)};
			cuuint64_t gstr[1] = {(cuuint64_t)w * 8};
			cuuint32_t box[2] = {(cuuint32_t)w, (cuuint32_t)boxes[b]};
			cuuint32_t es[2] = {1, 1};
			CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, w == 16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
			                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
			printf("width %d box rows %d: encode rc %d\n", w, boxes[b], (int)r);
			if (r) continue;
			for (int e = 0; e < 2; e++) {
				const uint32_t expect = (uint32_t)(4 * w * 8) * (e == 0 ? 1 : boxes[b]);
				if (e == 1 && boxes[b] == 1) continue;
				cudaMemset(st, 0, 4);
				probe<<<1, 128, 8192>>>(map, 5, 17, 2, 900, expect, out, st);
				cudaError_t ce = cudaDeviceSynchronize();
				int hs = 0;
				std::vector<double> ho(512);
				cudaMemcpy(&hs, st, 4, cudaMemcpyDeviceToHost);
				cudaMemcpy(ho.data(), out, 512 * 8, cudaMemcpyDeviceToHost);
				printf("  expect_tx %u: sync %s, barrier %s\n", expect, cudaGetErrorString(ce), hs == 1 ? "completed" : (hs == -1 ? "TIMED OUT" : "?"));
				if (ce != cudaSuccess) return 1;
				// where did the values land: print for each 16-byte chunk of the first 1024 bytes the (row, col) it holds
				for (int ch = 0; ch < 64; ch++) {
					const double v = ho[(size_t)ch * 2];
					if (v < 0) continue;
					const int row = (int)(v / 100.0), col = (int)(v - row * 100.0);
					printf("    smem chunk %2d (byte %4d): row %3d col %2d\n", ch, ch * 16, row, col);
				}
			}
		}
	}
	return 0;
}
