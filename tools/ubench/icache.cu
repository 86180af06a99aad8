// Instruction-cache probe: a loop of BODY_KB of straight-line code, 24 warps per SM
// (one 768-thread CTA on every SM), warps de-phased; cycles per warp-instruction
// against the size of the loop body.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned u32;
#define I1(x) asm volatile("lop3.b32 %0, %0, %1, 0x5c5c5c5c, 0x96;" : "+r"(x) : "r"(s));
#define I2(a,b) asm volatile("mad.lo.u32 %0, %0, %1, 0x01010101;" : "+r"(a) : "r"(b));
// 8 instructions, 4 independent chains, ALU/FMA mixed
#define B8 I1(x0) I2(x1,one) I1(x2) I2(x3,one) I1(x0) I2(x1,one) I1(x2) I2(x3,one)
#define B64 B8 B8 B8 B8 B8 B8 B8 B8
#define B512 B64 B64 B64 B64 B64 B64 B64 B64
#define KB1 B64
#define KB4 KB1 KB1 KB1 KB1
#define KB8 KB4 KB4
#define KB16 KB8 KB8
#define KB32 KB16 KB16
#define KB64 KB32 KB32

template <int KB> __device__ __forceinline__ void body(u32 &x0, u32 &x1, u32 &x2, u32 &x3, u32 s, u32 one);
#define DEF(N, CODE) template <> __device__ __forceinline__ void body<N>(u32 &x0, u32 &x1, u32 &x2, u32 &x3, u32 s, u32 one) { CODE }
DEF(4, KB4) DEF(8, KB8) DEF(12, KB8 KB4) DEF(16, KB16) DEF(20, KB16 KB4) DEF(24, KB16 KB8) DEF(28, KB16 KB8 KB4)
DEF(32, KB32) DEF(40, KB32 KB8) DEF(48, KB32 KB16) DEF(64, KB64) DEF(96, KB64 KB32) DEF(128, KB64 KB64)

template <int KB>
__global__ void __launch_bounds__(768, 1) k(u32 *out, u32 one, u32 s, int iters, int dephase, long long *cyc)
{
	u32 x0 = s + threadIdx.x, x1 = x0 * 3, x2 = x0 * 5, x3 = x0 * 7;
	const int w = threadIdx.x >> 5;
	// de-phase: warp w idles w * dephase cycles first
	if (dephase) {
		long long t = clock64();
		while (clock64() - t < (long long)w * dephase) { }
	}
	long long t0 = clock64();
#pragma unroll 1
	for (int it = 0; it < iters; it++)
		body<KB>(x0, x1, x2, x3, s, one);
	long long t1 = clock64();
	out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3;
	if (threadIdx.x == 0 && blockIdx.x == 0)
		*cyc = t1 - t0;
}

template <int KB> void run(u32 *out, long long *cyc, int dephase)
{
	const int ninstr = KB * 64;			// 16 B each
	const int iters = (4 << 20) / ninstr;		// ~4M instructions per warp
	k<KB><<<148, 768>>>(out, 1, 12345, 8, dephase, cyc);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaEventRecord(e0);
	k<KB><<<148, 768>>>(out, 1, 12345, iters, dephase, cyc);
	cudaEventRecord(e1);
	cudaDeviceSynchronize();
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
	double wi = 6.0 * ninstr * iters;		// per SMSP
	printf("body %3d KB  dephase %5d: %.3f warp-instr/cycle/SMSP (warp 0: %lld cycles, kernel %.2f ms)\n", KB, dephase, wi / c, c, ms);
}

int main()
{
	u32 *out; long long *cyc;
	cudaMalloc(&out, 148 * 768 * 4); cudaMalloc(&cyc, 8);
	for (int dp : { 0, 3000 }) {
		run<4>(out, cyc, dp); run<8>(out, cyc, dp); run<12>(out, cyc, dp); run<16>(out, cyc, dp);
		run<20>(out, cyc, dp); run<24>(out, cyc, dp); run<28>(out, cyc, dp); run<32>(out, cyc, dp);
		run<40>(out, cyc, dp); run<48>(out, cyc, dp); run<64>(out, cyc, dp); run<96>(out, cyc, dp); run<128>(out, cyc, dp);
	}
	return 0;
}
