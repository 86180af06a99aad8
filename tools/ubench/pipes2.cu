// Second probe: compare / select / predicate-producing ops, LDS, XU ops.
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
	__shared__ u32 sm[2048];
	sm[threadIdx.x] = threadIdx.x; sm[threadIdx.x + 1024] = 1;
	__syncthreads();
	const u32 sa = (u32)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 4;
	const u32 sa16 = (u32)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 16;
	long long t0 = clock64();
#pragma unroll 1
	for (int it = 0; it < ITERS; it++) {
#pragma unroll
		for (int i = 0; i < CHAINS; i++) {
			if (MODE == 0)		/* ISETP + predicated VIADD */
				asm volatile("{ .reg .pred p; setp.gt.u32 p, %0, %1; @p add.u32 %0, %0, 0x01010101; }" : "+r"(x[i]) : "r"(seed));
			else if (MODE == 1)	/* selp only (pred constant-ish) */
				asm volatile("{ .reg .pred p; setp.gt.u32 p, %1, 5; selp.u32 %0, %0, %2, p; }" : "+r"(x[i]) : "r"(seed), "r"(one));
			else if (MODE == 2)	/* lop3 with predicate result: and.b32 then setp.ne fused? */
				asm volatile("{ .reg .pred p; .reg .u32 t; and.b32 t, %0, 0x80808080; setp.ne.u32 p, t, 0; @p add.u32 %0, %0, 0x01010101; }" : "+r"(x[i]));
			else if (MODE == 3)	/* LDS.32 conflict free */
				asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x[i]) : "r"(sa + ((i * 128) & 4095)));
			else if (MODE == 4) {	/* LDS.128 conflict free */
				u32 a, b, c;
				asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x[i]), "=r"(a), "=r"(b), "=r"(c) : "r"(sa16 + ((i * 512) & 4095)));
			}
			else if (MODE == 5)	/* popc */
				asm volatile("popc.b32 %0, %0;" : "+r"(x[i]));
			else if (MODE == 6)	/* brev+bfind style ffs: brev then clz */
				asm volatile("{ .reg .u32 t; brev.b32 t, %0; clz.b32 %0, t; }" : "+r"(x[i]));
			else if (MODE == 7)	/* vote.ballot */
				asm volatile("{ .reg .pred p; setp.ne.u32 p, %0, 0; vote.sync.ballot.b32 %0, p, 0xffffffff; }" : "+r"(x[i]));
			else if (MODE == 8)	/* shfl */
				asm volatile("shfl.sync.idx.b32 %0, %0, 3, 0x1f, 0xffffffff;" : "+r"(x[i]));
			else if (MODE == 9)	/* min.u32 (VIMNMX) */
				asm volatile("min.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(seed));
			else if (MODE == 10)	/* shl imm */
				asm volatile("shl.b32 %0, %0, 3;" : "+r"(x[i]));
			else if (MODE == 11)	/* shr imm */
				asm volatile("shr.u32 %0, %0, 3;" : "+r"(x[i]));
			else if (MODE == 12)	/* bfe */
				asm volatile("bfe.u32 %0, %0, 12, 12;" : "+r"(x[i]));
			else if (MODE == 13)	/* add3 reg reg */
				asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(seed));
			else if (MODE == 14)	/* LDS.U8 */
				asm volatile("ld.shared.u8 %0, [%1];" : "=r"(x[i]) : "r"(sa + ((i * 128) & 4095)));
			else if (MODE == 15) {	/* LOP3 + LDS32 2:1 conflict free */
				if (i % 3 == 2)
					asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x[i]) : "r"(sa + ((i * 128) & 4095)));
				else
					asm volatile("lop3.b32 %0, %0, %1, 0x5c5c5c5c, 0x96;" : "+r"(x[i]) : "r"(seed));
			}
			else if (MODE == 16)	/* setp + bra (taken, uniform) */
				asm volatile("{ .reg .pred p; setp.eq.u32 p, %0, 0x12345; @p bra L%=; add.u32 %0, %0, 1; L%=: }" : "+r"(x[i]));
			else if (MODE == 17)	/* dp4a */
				asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(one), "r"(seed));
			else if (MODE == 18)	/* vabsdiff4 */
				asm volatile("vabsdiff4.u32.u32.u32.add %0, %0, %1, %2;" : "+r"(x[i]) : "r"(one), "r"(seed));
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
	double wi = (NT / 128.0) * CHAINS * ITERS;
	printf("%-28s %8lld cycles  %.3f asm-stmts/cycle/SMSP\n", name, c, wi / c);
}
int main()
{
	u32 *out; long long *cyc;
	cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
	NT = 768;
	run<0>("ISETP + @p VIADD", out, cyc);
	run<1>("ISETP(inv) + SEL", out, cyc);
	run<2>("AND+ISETP.NE + @p VIADD", out, cyc);
	run<3>("LDS.32", out, cyc);
	run<4>("LDS.128", out, cyc);
	run<14>("LDS.U8", out, cyc);
	run<15>("LOP3+LOP3+LDS", out, cyc);
	run<5>("POPC", out, cyc);
	run<6>("BREV+FLO", out, cyc);
	run<7>("ISETP+VOTE", out, cyc);
	run<8>("SHFL", out, cyc);
	run<9>("VIMNMX", out, cyc);
	run<10>("SHL imm", out, cyc);
	run<11>("SHR imm", out, cyc);
	run<12>("BFE", out, cyc);
	run<13>("IADD r,r", out, cyc);
	run<16>("ISETP+BRA(not taken)+VIADD", out, cyc);
	run<17>("DP4A", out, cyc);
	run<18>("VABSDIFF4", out, cyc);
	return 0;
}
