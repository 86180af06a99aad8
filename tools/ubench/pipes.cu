// Issue-rate probe for the integer pipes of sm_100a: how many warp-instructions
// per cycle per SM sub-partition (SMSP) for LOP3 / IADD3 / IMAD / SHF / mixes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned u32;
#define CHAINS 8
#define ITERS 512

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(u32 *out, u32 one, u32 seed, long long *cyc)
{
	u32 x[CHAINS];
	for (int i = 0; i < CHAINS; i++)
		x[i] = seed + threadIdx.x * 17 + i;
	__shared__ u32 sm[1024];
	sm[threadIdx.x] = threadIdx.x;
	__syncthreads();
	long long t0 = clock64();
#pragma unroll 1
	for (int it = 0; it < ITERS; it++) {
#pragma unroll
		for (int i = 0; i < CHAINS; i++) {
			if (MODE == 0)		/* LOP3 reg,reg,imm */
				asm volatile("lop3.b32 %0, %0, %1, 0x5c5c5c5c, 0x96;" : "+r"(x[i]) : "r"(seed));
			else if (MODE == 1)	/* add reg,imm */
				asm volatile("add.u32 %0, %0, 0x01010101;" : "+r"(x[i]));
			else if (MODE == 2)	/* mad reg,reg,imm (FMA pipe) */
				asm volatile("mad.lo.u32 %0, %0, %1, 0x01010101;" : "+r"(x[i]) : "r"(one));
			else if (MODE == 3) {	/* LOP3 + IMAD alternating */
				if (i & 1)
					asm volatile("mad.lo.u32 %0, %0, %1, 0x01010101;" : "+r"(x[i]) : "r"(one));
				else
					asm volatile("lop3.b32 %0, %0, %1, 0x5c5c5c5c, 0x96;" : "+r"(x[i]) : "r"(seed));
			} else if (MODE == 4)	/* funnel shift, register amount */
				asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(seed), "r"(one));
			else if (MODE == 5) {	/* LOP3 + add alternating (both ALU?) */
				if (i & 1)
					asm volatile("add.u32 %0, %0, 0x01010101;" : "+r"(x[i]));
				else
					asm volatile("lop3.b32 %0, %0, %1, 0x5c5c5c5c, 0x96;" : "+r"(x[i]) : "r"(seed));
			} else if (MODE == 6) {	/* 2 LOP3 : 1 IMAD */
				if (i % 3 == 2)
					asm volatile("mad.lo.u32 %0, %0, %1, 0x01010101;" : "+r"(x[i]) : "r"(one));
				else
					asm volatile("lop3.b32 %0, %0, %1, 0x5c5c5c5c, 0x96;" : "+r"(x[i]) : "r"(seed));
			} else if (MODE == 7) {	/* LOP3 + LDS alternating */
				if (i & 1)
					asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x[i]) : "r"((u32)__cvta_generic_to_shared(sm) + ((x[i-1] & 1023) << 2)));
				else
					asm volatile("lop3.b32 %0, %0, %1, 0x5c5c5c5c, 0x96;" : "+r"(x[i]) : "r"(seed));
			} else if (MODE == 8) {	/* setp + selp */
				asm volatile("{ .reg .pred p; setp.ne.u32 p, %0, %1; selp.u32 %0, %0, %2, p; }" : "+r"(x[i]) : "r"(seed), "r"(one));
			} else if (MODE == 9)	/* LOP3 reg,reg,reg */
				asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(seed), "r"(one));
			else if (MODE == 10)	/* prmt */
				asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(seed), "r"(one));
			else if (MODE == 11) {	/* sub via mad with -1 multiplier imm? mul.lo by imm */
				asm volatile("mad.lo.u32 %0, %0, 3, %1;" : "+r"(x[i]) : "r"(seed));
			} else if (MODE == 13) {	/* FFMA r,r,r */
				asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(seed), "r"(one));
			} else if (MODE == 14) {	/* FFMA r,imm,r */
				asm volatile("fma.rn.f32 %0, %0, 0f3f800001, %1;" : "+r"(x[i]) : "r"(seed));
			} else if (MODE == 15) {	/* LOP3 + FFMA */
				if (i & 1)
					asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(seed), "r"(one));
				else
					asm volatile("lop3.b32 %0, %0, %1, 0x5c5c5c5c, 0x96;" : "+r"(x[i]) : "r"(seed));
			} else if (MODE == 16) {	/* LOP3 + VIADD + IMAD */
				if (i % 4 == 0 || i % 4 == 2)
					asm volatile("lop3.b32 %0, %0, %1, 0x5c5c5c5c, 0x96;" : "+r"(x[i]) : "r"(seed));
				else if (i % 4 == 1)
					asm volatile("add.u32 %0, %0, 0x01010101;" : "+r"(x[i]));
				else
					asm volatile("mad.lo.u32 %0, %0, %1, 0x01010101;" : "+r"(x[i]) : "r"(one));
			} else if (MODE == 17) {	/* LOP3 imm-only source: r, imm (2-input) */
				asm volatile("xor.b32 %0, %0, 0x5c5c5c5c;" : "+r"(x[i]));
			} else if (MODE == 18) {	/* xor imm + VIADD alternating */
				if (i & 1)
					asm volatile("add.u32 %0, %0, 0x01010101;" : "+r"(x[i]));
				else
					asm volatile("xor.b32 %0, %0, 0x5c5c5c5c;" : "+r"(x[i]));
			} else if (MODE == 12) { /* 1 LOP3 : 1 IMAD : 1 LDS */
				if (i % 3 == 0)
					asm volatile("lop3.b32 %0, %0, %1, 0x5c5c5c5c, 0x96;" : "+r"(x[i]) : "r"(seed));
				else if (i % 3 == 1)
					asm volatile("mad.lo.u32 %0, %0, %1, 0x01010101;" : "+r"(x[i]) : "r"(one));
				else
					asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x[i]) : "r"((u32)__cvta_generic_to_shared(sm) + ((x[i-1] & 1023) << 2)));
			}
		}
	}
	long long t1 = clock64();
	u32 s = 0;
	for (int i = 0; i < CHAINS; i++)
		s ^= x[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0 && blockIdx.x == 0)
		*cyc = t1 - t0;
}

static int NT = 768;
template <int MODE> void run(const char *name, u32 *out, long long *cyc)
{
	k<MODE><<<148, NT>>>(out, 1, 12345, cyc);
	k<MODE><<<148, NT>>>(out, 1, 12345, cyc);
	cudaDeviceSynchronize();
	long long c;
	cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
	// per SMSP: 6 warps x CHAINS x ITERS warp-instructions (+ loop overhead)
	double wi = (NT / 128.0) * CHAINS * ITERS;
	printf("%-28s %8lld cycles  %.3f warp-instr/cycle/SMSP\n", name, c, wi / c);
}

int main()
{
	u32 *out; long long *cyc;
	cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
	for (int nt : { 128, 256, 512, 768, 1024 }) {
	NT = nt;
	printf("---- %d threads per SM (%d warps per SMSP)\n", nt, nt / 128);
	run<0>("LOP3 r,r,imm", out, cyc);
	run<9>("LOP3 r,r,r", out, cyc);
	run<1>("IADD r,imm", out, cyc);
	run<2>("IMAD r,r,imm", out, cyc);
	run<11>("IMAD r,imm,r", out, cyc);
	run<4>("SHF r,r,r", out, cyc);
	run<10>("PRMT r,r,r", out, cyc);
	run<8>("ISETP+SEL", out, cyc);
	run<5>("LOP3+IADD 1:1", out, cyc);
	run<3>("LOP3+IMAD 1:1", out, cyc);
	run<6>("LOP3+IMAD 2:1", out, cyc);
	run<7>("LOP3+LDS 1:1", out, cyc);
	run<12>("LOP3+IMAD+LDS 1:1:1", out, cyc);
	run<13>("FFMA r,r,r", out, cyc);
	run<14>("FFMA r,imm,r", out, cyc);
	run<15>("LOP3+FFMA 1:1", out, cyc);
	run<16>("LOP3+VIADD+LOP3+IMAD", out, cyc);
	run<17>("XOR r,imm", out, cyc);
	run<18>("XOR imm + VIADD 1:1", out, cyc);
	}
	return 0;
}
