// Persistent, warp-specialised, TMA-fed RK-stage kernel ("v4"): one_rk_step_M1/_M2 (cfd_v0.cpp:2530/1897) +
// prepare_for_RKstep (1339) on the record layout (kernels.cuh, Rec<D>).
//
// Why: the round-1 tile kernel (k_tile_stage) spends half of its cycles on a serial latency chain per tile -- tile
// descriptor -> halo cell ids -> 6 624 eight-byte cp.async element copies -> first face constants -- with only two CTAs
// per SM to hide it (profiles/r2a_stage128_stall_table.txt: 37 % long-scoreboard + 17 % barrier stalls).  Here one CTA per
// SM lives for the whole launch and splits into
//   * a PRODUCER warp that runs ahead of the arithmetic: per tile it fills the next free slot of a shared-memory ring with
//       - the Q and V records of the tile's own cells: two TMA tensor loads (cp.async.bulk.tensor.2d, 128B/64B/32B
//         hardware swizzle), completion counted in bytes on the slot's `full` mbarrier;
//       - the records of the halo cells: 16-byte cp.async copies (a halo cell is 12 of them instead of 23 element copies
//         from 23 different sectors), written with the same swizzle, completion signalled on the same mbarrier
//         (cp.async.mbarrier.arrive.noinc);
//     descriptors and halo ids are fetched one tile ahead, so the chain above is off the consumers' critical path;
//   * two CONSUMER groups of 256 threads that take alternate tiles: wait on `full`, evaluate every face of the tile once
//     (both sides read with conflict-free 16-byte LDS from the swizzled records: 24 loads per face instead of 46), ordered
//     gather + sponge + RK update per cell, derived values, and hand the tile's new Q records to the TMA engine as ONE bulk
//     store (cp.async.bulk.global.shared::cta); then release the slot on its `empty` mbarrier.
// The arithmetic is device_math.cuh's, in the same order: results are bit-identical to k_tile_stage and to the reference.
#pragma once
#include <cuda.h>

#include "tile_kernels.cuh"

namespace lfm {

// ---- PTX wrappers ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// The suspend-time hint lets the hardware park the thread until the phase completes (or the hint, in nanoseconds, runs out)
// instead of returning at once: without it twenty warps polling their barriers issued a sixth of all instructions of the
// kernel (profiles/r2j_*), taken from the warps that had arithmetic to do.
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
	uint32_t ok;
	asm volatile(
	    "{\n"
	    ".reg .pred p;\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
	    "selp.u32 %0, 1, 0, p;\n"
	    "}"
	    : "=r"(ok)
	    : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
	    : "memory");
	return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
	while (!mbar_try_wait(bar, parity)) {
	}
}
// this thread's earlier cp.async copies arrive on the mbarrier when they have landed (the arrival is pre-counted at init)
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) { asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }
// (no "memory" clobber: the producers never touch the destination with ordinary accesses, and volatile asm statements keep
// their order among themselves -- the mbarrier wait before and the arrive after; this lets the compiler batch the loads of
// the halo ids of an unrolled copy loop instead of serialising id load -> address -> copy per element)
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src)); }
// TMA: box of the 2D tensor map at (c0 = value index inside the record, c1 = first cell) -> shared memory, bytes counted on bar
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst_smem), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
	             : "memory");
}
// TMA gather (sm_100 tile::gather4): FOUR rows of the 2D tensor, named by their indices, land in four consecutive row slots
// of shared memory (swizzled by address like a box); the tensor map's box is one row.  Probed on B200 with
// scripts/probe/gather4_probe.cu: a box of one row is required (a box of four rows is an illegal instruction), one
// instruction delivers 4 x row bytes.
__device__ __forceinline__ void tma_gather4(uint32_t dst_smem, const CUtensorMap* map, const int4& rows, uint64_t* bar) {
	asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst_smem), "l"(map), "r"(0), "r"(rows.x),
	             "r"(rows.y), "r"(rows.z), "r"(rows.w), "r"(smem_u32(bar))
	             : "memory");
}
// TMA: contiguous shared memory -> global memory (sizes and addresses multiples of 16 bytes)
__device__ __forceinline__ void bulk_store(void* dst_global, uint32_t src_smem, uint32_t bytes) {
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_global), "r"(src_smem), "r"(bytes) : "memory");
}
// TMA: ask the L2 to fetch a contiguous piece of global memory (address and size multiples of 16 bytes); nothing waits for it
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) { asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes)); }
// one cache line into the L2, per lane (LSU instruction: no uniform-datapath serialisation, nothing waits for it)
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// ---- swizzled records in shared memory -----------------------------------------------------------------------------------
// A region holds records of B = 32, 64 or 128 bytes at consecutive slots, laid out as the TMA engine writes a box of a
// tensor map with CU_TENSOR_MAP_SWIZZLE_{32,64,128}B: the 16-byte chunk index (address bits 4..6) is XORed with address
// bits 7..9 (128B; two bits for 64B, one for 32B).  Regions start on 1024-byte boundaries, so offsets stand for addresses.
// Eight consecutive slots then cover all 32 banks with any one chunk: 16-byte loads of consecutive cells never conflict.
__host__ __device__ constexpr uint32_t swz_mask(int record_bytes) { return (uint32_t)(record_bytes - 16) & 0x70u; }
__device__ __forceinline__ uint32_t swz(uint32_t off, uint32_t mask) { return off ^ ((off >> 3) & mask); }

template <class R> struct Chunk16;
template <> struct Chunk16<double> {
	using T = double2;
	static constexpr int N = 2;
	static __device__ __forceinline__ void unpack(const T& t, double* d) {
		d[0] = t.x;
		d[1] = t.y;
	}
	static __device__ __forceinline__ T pack(const double* d) { return make_double2(d[0], d[1]); }
	static __device__ __forceinline__ double get(const T& t, int i) { return i == 0 ? t.x : t.y; }
};
template <> struct Chunk16<float> {
	using T = float4;
	static constexpr int N = 4;
	static __device__ __forceinline__ void unpack(const T& t, float* d) {
		d[0] = t.x;
		d[1] = t.y;
		d[2] = t.z;
		d[3] = t.w;
	}
	static __device__ __forceinline__ T pack(const float* d) { return make_float4(d[0], d[1], d[2], d[3]); }
	static __device__ __forceinline__ float get(const T& t, int i) { return i == 0 ? t.x : (i == 1 ? t.y : (i == 2 ? t.z : t.w)); }
};

// one side of a face in registers, filled by 16-byte loads from the swizzled records of a ring slot
template <class R, int D> struct RecSide {
	using RC = Rec<D>;
	using C16 = Chunk16<R>;
	static constexpr bool kHasTrace = (D == 3);
	static constexpr int QB = RC::QW * (int)sizeof(R), VB = RC::VW * (int)sizeof(R);
	typename C16::T qc[QB / 16];   // the chunks as loaded: the accessors name a half (quarter) of one, no value is moved
	typename C16::T vc[VB / 16];
	__device__ __forceinline__ void load_q(const unsigned char* Qs, int slot) {
		const uint32_t p = swz((uint32_t)slot * QB, swz_mask(QB));
#pragma unroll
		for (int j = 0; j < QB / 16; j++) qc[j] = *reinterpret_cast<const typename C16::T*>(Qs + (p ^ (uint32_t)(j << 4)));
	}
	__device__ __forceinline__ void load_v(const unsigned char* Vs, int slot) {
		const uint32_t p = swz((uint32_t)slot * VB, swz_mask(VB));
#pragma unroll
		for (int j = 0; j < VB / 16; j++) vc[j] = *reinterpret_cast<const typename C16::T*>(Vs + (p ^ (uint32_t)(j << 4)));
	}
	__device__ __forceinline__ R qval(int k) const { return C16::get(qc[k / C16::N], k % C16::N); }
	__device__ __forceinline__ R vval(int k) const { return C16::get(vc[k / C16::N], k % C16::N); }
	__device__ __forceinline__ R q(int k) const { return qval(k); }
	__device__ __forceinline__ R rho_inv() const { return qval(RC::RHO_INV); }
	__device__ __forceinline__ R Rpsi() const { return qval(RC::RPSI); }
	__device__ __forceinline__ R aux() const { return qval(RC::AUX); }
	__device__ __forceinline__ R dudx(int a, int b) const { return vval(RC::DUDX + a * D + b); }
	__device__ __forceinline__ R dTdx(int a) const { return vval(RC::DTDX + a); }
	__device__ __forceinline__ R sigmaU(int a) const { return vval(RC::SIGMAU + a); }
	__device__ __forceinline__ R trace_neg() const { return vval(D == 3 ? RC::TR : 0); }
	__device__ __forceinline__ R tauMC(int, int) const { return R(0); }   // laminar closure only
};

// ---- producer warps ---------------------------------------------------------------------------------------------------
// Each of the four producer warps is a complete producer for the tiles of "its" ring slots (tile i of the CTA uses slot i % NS
// and belongs to warp (i % NS) % 4, so the uses of one slot are filled in order by one warp -- a second warp could reach the
// slot two phases early, which a parity wait cannot tell from "released"): a serial loop per tile costs about two thousand cycles of dependent latencies (barrier tests, shared-memory
// round trips, descriptor fetch), which one warp alone cannot hide when a tile's arithmetic is shorter than that (the
// gradient kernel); four independent loops can.  Per tile a warp
//   * has the tile's descriptor fields in registers (read from global memory two of its own iterations earlier) and the tile's
//     halo cell ids in its private two-deep ring in shared memory (a TMA bulk copy issued one iteration earlier, completion
//     on the ring's own mbarrier: the plan pads every halo list to a 16-byte boundary);
//   * waits for the ring slot to be released, then issues the TMA loads of the own cells and the 16-byte cp.async copies of
//     the halo records, whose completion arrives on the slot's `full` mbarrier.
constexpr int kProducerWarps = 4;
struct TileMeta {
	int c0, nh, halo_off, f_off, nf;
};
__device__ __forceinline__ TileMeta meta_of(const TileDesc* d) {
	const int4 a = *reinterpret_cast<const int4*>(d);              // c0, nt, halo_off, nh
	const int4 b = *reinterpret_cast<const int4*>(&d->f_off);      // f_off, nfo, ninc, hb
	return TileMeta{a.x, a.w, a.z, b.x, b.y + b.z};
}
// lane 0: start the bulk copy of a tile's halo ids into an id ring slot (nothing to copy: a plain arrival completes the phase)
__device__ __forceinline__ void ids_fetch(const TileMeta& mt, int* ring_ids, uint64_t* bar, const int* __restrict__ halo_cell) {
	const uint32_t bytes = ((uint32_t)mt.nh + 3u) / 4u * 16u;
	if (bytes) {
		mbar_arrive_expect_tx(bar, bytes);
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(ring_ids)), "l"(halo_cell + mt.halo_off), "r"(bytes), "r"(smem_u32(bar))
		             : "memory");
	} else {
		mbar_arrive(bar);
	}
}

// what the host passes about the ring (all byte offsets from the 1024-aligned start of dynamic shared memory)
struct PipeGeom {
	int n_slots;            // ring depth
	int box_cells;          // rows of the TMA box = first halo slot
	int smax;               // slots per region (box_cells + largest halo)
	int fmax;               // faces per tile (stride of the flux rows)
	int hmax;               // largest halo
	uint32_t off_bar;       // full[n_slots], empty[n_slots]
	uint32_t off_ids;       // [kProducerWarps][2][hmax4] int: the producer warps' halo id rings
	uint32_t off_meta;      // [kProducerWarps][2] mbarriers of the id rings
	uint32_t off_fl;        // [2][NQ][fmax] R face fluxes
	uint32_t off_slot;      // first ring slot
	uint32_t slot_bytes;    // Q region + V region (each a multiple of 1024)
	uint32_t q_bytes;       // Q region
	int pf_dist;            // L2 prefetch distance in tiles of this CTA (0: none): own-cell records and face-table slices of the tile
	                        // pf_dist fills ahead are requested into the L2, so that a fill costs an L2 round trip, not a DRAM one
	int dbg;                // timing experiments only (wrong results): bit 0 no halo copies, bit 1 no own-cell loads, bit 2 consumers skip
	                        // the arithmetic (wait for the slot, release it): LFMGPU_PIPE_DBG
	int coop;               // 1: the four producer warps share every fill (a quarter of the halo gathers each); 0: one warp per ring slot
	int wstore;             // 1: every consumer warp stores the records of its own cells and releases the slot for itself (no group barrier
	                        // behind phase C; the flux rows are guarded by an mbarrier instead); 0: one store per tile behind a group barrier
	int direct;             // 1: the finished records go from registers straight to HBM (two 16-byte stores per lane = whole sectors) instead of
	                        // through a staging area and a bulk store; slots and flux rows are released per warp as with wstore
	int pf_face;            // 1: the consumers ask for the face constants of their second round (L2 hits after the producers' prefetch) into
	                        // the L1 before they start the first: the round then starts on L1 hits instead of an L2 round trip
	int pf_cell;            // 1: the L2 prefetch also covers what phase C reads per cell (gather lists, accumulators, volumes, sponge)
};

constexpr int kPipeGroupThreads = 256;                      // one consumer group (two warpgroups)
constexpr int kPipeProducerThreads = 128;                   // one warpgroup: setmaxnreg is a warpgroup-wide instruction
constexpr int kPipeThreads = 2 * kPipeGroupThreads + kPipeProducerThreads;   // 20 warps: five per scheduler, 96 registers each at launch
constexpr int kPipeMaxHalo = 256;                           // largest halo the id ring is laid out for
// Register budget: five warps per scheduler cap the launch at 96 registers per thread (640 x 96 = 61 440: what the CTA owns
// for its whole life -- setmaxnreg only moves registers between its warps).  The producer warpgroup gives back all but 32 per
// thread and the consumers grow to 112: 128 x 32 + 512 x 112 = 61 440.
constexpr int kPipeProducerRegs = 32, kPipeConsumerRegs = 112;
static_assert(kPipeProducerThreads * kPipeProducerRegs + 2 * kPipeGroupThreads * kPipeConsumerRegs <= kPipeThreads * 96, "setmaxnreg.inc would wait for registers the CTA does not own");

// Phase C for components [I0, I1) of tile cell lc on the record layout (gather_update of tile_kernels.cuh): dq = A_k dq +
// sum_faces(+-dt rhs / V) + sponge, q_new = q + B_k dq (prepare_for_RKstep's scaling, the scatter of cfd_v0.cpp:2792-2804 as
// an ordered gather, sponge 2810-2814, update 2819-2822).  The cell's conservatives are read from its Q record in the ring
// slot; the new ones are returned in qnew[I0 .. I1).  `x -= y` is evaluated as `x += (-y)`: bit-identical, branch-free.
// gather_inputs issues the global loads; the caller puts the group barrier that closes phase B between the two (the halves
// of a warp run different instantiations: a barrier inside them would be reached divergently), so the loads travel during
// the wait.
// NC components starting at the run-time component cb (both lanes of a cell run the same code; only addresses depend on cb)
template <class R, int D, int NC> struct CellIn {   // what phase C reads from global memory for one cell
	int e[kMaxSlots];
	R dq[NC], vinv, sg;
};
template <class R, int D, int NC> __device__ __forceinline__ void gather_inputs(const DevMesh<R>& m, const TileView<R>& tv, int c, int cb, bool active, int first, CellIn<R, D, NC>& in) {
#pragma unroll
	for (int s = 0; s < kMaxSlots; s++) in.e[s] = 0;
#pragma unroll
	for (int i = 0; i < NC; i++) in.dq[i] = R(0);
	in.vinv = in.sg = R(0);
	if (active) {
#pragma unroll
		for (int s = 0; s < kMaxSlots; s++)
			if (s < m.F) in.e[s] = (int)tv.csr_local[(size_t)s * m.n_cells + c];
		if (!first) {
#pragma unroll
			for (int i = 0; i < NC; i++) in.dq[i] = m.dq[(size_t)(cb + i) * m.n_cells + c];
		}
		in.vinv = m.vol_inv[c];
		in.sg = m.sigma[c];
	}
}
__device__ __forceinline__ double flip_sign(double v, bool neg) { return __hiloint2double(__double2hiint(v) ^ (neg ? (int)0x80000000u : 0), __double2loint(v)); }
__device__ __forceinline__ float flip_sign(float v, bool neg) { return __int_as_float(__float_as_int(v) ^ (neg ? (int)0x80000000u : 0)); }
template <class R, int D, int NC>
__device__ __forceinline__ void gather_update_rec(const DevMesh<R>& m, CellIn<R, D, NC>& in, int c, int cb, const unsigned char* Qs, const R* fl, int fmax, int lc, R dt, R Ak, R Bk, int res, R* qnew) {
	constexpr int NQ = D + 2, QB = Rec<D>::QW * (int)sizeof(R);
	R RES[NC];
#pragma unroll
	for (int i = 0; i < NC; i++) {
		in.dq[i] *= Ak;
		RES[i] = R(0);
	}
	const R* flc = fl + (size_t)cb * fmax;
#pragma unroll
	for (int s = 0; s < kMaxSlots; s++) {
		// (a cell's list is dense: entries behind the first 0 are 0 too.  No early exit, so that the loads of all slots can be
		// issued together; an empty slot contributes nothing -- not even a +0 that could turn a -0 sum into +0)
		const bool used = in.e[s] != 0, own = in.e[s] > 0;
		const int lfc = used ? (own ? in.e[s] : -in.e[s]) - 1 : 0;
		R v[NC];
#pragma unroll
		for (int i = 0; i < NC; i++) v[i] = flc[i * fmax + lfc];
		if (used) {
#pragma unroll
			for (int i = 0; i < NC; i++) {
				const R rr = flip_sign(v[i], !own);   // own ? v : -v as one integer XOR on the sign bit (a DADD and two selects otherwise)
				if (res) RES[i] += rr;
				in.dq[i] += dt * rr * in.vinv;
			}
		}
	}
	const uint32_t p = swz((uint32_t)lc * QB, swz_mask(QB));
#pragma unroll
	for (int i = 0; i < NC; i++) {
		const int k = cb + i;   // component
		const R cqi = *reinterpret_cast<const R*>(Qs + (p ^ (uint32_t)(((k * (int)sizeof(R)) >> 4) << 4)) + ((k * (int)sizeof(R)) & 15));
		const R target = k == 0 ? m.k.rhoInf : (k == NQ - 1 ? m.k.rhoEInf : m.k.rhoUInf[(k > 0 && k < NQ - 1) ? k - 1 : 0]);
		in.dq[i] += dt * in.sg * (target - cqi);
		m.dq[(size_t)k * m.n_cells + c] = in.dq[i];
		qnew[i] = cqi + Bk * in.dq[i];
		if (res) m.RES[(size_t)k * m.n_cells + c] = RES[i];
	}
}

// exchange of a value between the two lanes of a cell (lanes l and l ^ 16)
__device__ __forceinline__ double pair_swap(double v) { return __shfl_xor_sync(0xffffffffu, v, 16); }
__device__ __forceinline__ float pair_swap(float v) { return __shfl_xor_sync(0xffffffffu, v, 16); }

template <class R, int D, int SCHEME>
__global__ void __launch_bounds__(kPipeThreads, 1)
    k_stage_pipe(DevMesh<R> m, TileView<R> tv, const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap gmap_q,
                 const __grid_constant__ CUtensorMap gmap_v, PipeGeom pg, const R* __restrict__ q, R* __restrict__ qn, int tile0,
                 int n_tiles, R dt, R Ak, R Bk, int first, int res) {
	using RC = Rec<D>;
	constexpr int NQ = D + 2, QB = RC::QW * (int)sizeof(R), VB = RC::VW * (int)sizeof(R), CQ = QB / 16;
	constexpr int GT = kPipeGroupThreads;
	extern __shared__ unsigned char smem_dyn[];
	unsigned char* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
	uint64_t* full = reinterpret_cast<uint64_t*>(sm + pg.off_bar);
	uint64_t* empty = full + pg.n_slots;
	const int NS = pg.n_slots, G = gridDim.x;
	const int lane = threadIdx.x & 31;

	if (threadIdx.x == 0) {
		for (int s = 0; s < NS; s++) {
			mbar_init(full + s, pg.coop ? kProducerWarps : 1);   // the expect_tx arrival of each producer warp that fills the slot; everything else is counted in bytes
			mbar_init(empty + s, (pg.wstore || pg.direct) ? GT / 32 : 1);   // one elected consumer thread (wstore, direct: one lane of every warp of the group)
		}
		for (int k = 0; k < 2; k++) mbar_init(full + 2 * pg.n_slots + k, GT / 32);   // flfree[group]: every warp of the group is done with the flux rows
		for (int k = 0; k < 2 * kProducerWarps; k++) mbar_init(reinterpret_cast<uint64_t*>(sm + pg.off_meta) + k, 1);   // the producers' id rings
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();

	if (threadIdx.x >= 2 * GT) {
		// ============================ producers (one warpgroup) ===================================================
		// Two roles, so that neither needs more than a handful of live registers (the producers run on 32; a spill costs an L2
		// round trip here, because shared memory leaves the L1 almost no capacity):
		//   warp 0     descriptors and halo ids one tile ahead -> shared memory, the TMA loads of the own cells, L2 prefetch;
		//   warps 1-3  the 16-byte cp.async copies of the halo records.
		asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kPipeProducerRegs));
		if ((int)blockIdx.x >= n_tiles) return;
		const int pwarp = (threadIdx.x - 2 * GT) >> 5;
		const int hpitch = (pg.hmax + 3) & ~3;
		int* idring = reinterpret_cast<int*>(sm + pg.off_ids) + pwarp * 2 * hpitch;
		uint64_t* idbar = reinterpret_cast<uint64_t*>(sm + pg.off_meta) + pwarp * 2;
		const unsigned char* qsrc = reinterpret_cast<const unsigned char*>(q);
		const unsigned char* vsrc = reinterpret_cast<const unsigned char*>(m.vis);
		const int my_tiles = (n_tiles - (int)blockIdx.x + G - 1) / G;
		const TileDesc* my_desc = tv.tiles + tile0 + blockIdx.x;   // tile i of this CTA: my_desc[i * G]
		// L2 prefetch of a tile some fills ahead: its own-cell records (two contiguous pieces), its slice of each face table and
		// (pf_cell) the per-cell inputs of phase C, which no copy stages: gather lists, RK accumulators, volumes, sponge
		auto prefetch_tile = [&](const TileDesc* d) {
			const TileMeta pm = meta_of(d);
			const int ac0 = pm.c0, af = pm.f_off;
			const uint32_t nf4 = (uint32_t)(pm.nf + 3) & ~3u;
			if (lane == 0) bulk_prefetch_l2(qsrc + (size_t)ac0 * QB, (uint32_t)pg.box_cells * QB);
			if (lane == 1) bulk_prefetch_l2(vsrc + (size_t)ac0 * VB, (uint32_t)pg.box_cells * VB);
			if (nf4) {
				if (lane == 2) bulk_prefetch_l2(tv.f_idx + af, nf4 * 4u);
				if (lane >= 3 && lane < 3 + D) bulk_prefetch_l2(tv.fS + (size_t)(lane - 3) * tv.T + af, nf4 * (uint32_t)sizeof(R));
				if (lane >= 3 + D && lane < 3 + 2 * D) bulk_prefetch_l2(tv.fK + (size_t)(lane - 3 - D) * tv.T + af, nf4 * (uint32_t)sizeof(R));
				if (lane == 3 + 2 * D) bulk_prefetch_l2(tv.fw + af, nf4 * (uint32_t)sizeof(R));
				if (lane == 4 + 2 * D) bulk_prefetch_l2(tv.fdm + af, nf4 * (uint32_t)sizeof(R));
				if (lane == 5 + 2 * D) bulk_prefetch_l2(tv.fdi + af, nf4 * (uint32_t)sizeof(R));
				if (SCHEME == 0 && lane == 6 + 2 * D) bulk_prefetch_l2(tv.fSmag + af, nf4 * (uint32_t)sizeof(R));
			}
			if (pg.pf_cell) {
				// one 128-byte line per lane and round; the first line of a row may start before the tile, the last one is clamped to
				// the last cell of the rank (a prefetch names an element that exists)
				constexpr int CPL = 128 / (int)sizeof(R);                       // cells per line, R rows
				const int nl = pg.box_cells / CPL + 1, nl16 = pg.box_cells / 64 + 1;   // lines per row (at most one round of the warp: box <= 256 cells)
				const int cr = min(ac0 + lane * CPL, m.n_cells - 1), c16 = min(ac0 + lane * 64, m.n_cells - 1);
				if (lane < nl) {
					prefetch_l2(m.vol_inv + cr);
					prefetch_l2(m.sigma + cr);
					if (!first) {
#pragma unroll
						for (int k = 0; k < NQ; k++) prefetch_l2(m.dq + (size_t)k * m.n_cells + cr);
					}
				}
				if (lane < nl16) {
#pragma unroll
					for (int sl = 0; sl < kMaxSlots; sl++)
						if (sl < m.F) prefetch_l2(tv.csr_local + (size_t)sl * m.n_cells + c16);
				}
			}
		};
		const int PD = pg.pf_dist;
		if (pg.coop) {
			// Every producer warp walks every tile of the CTA and issues a quarter of its fill.  A TMA gather names its rows in uniform
			// registers, so a warp issues the gathers of its lanes one after the other (about 15 instructions each): one warp alone
			// needs longer for the ~90 gathers of a fill than the consumers need for a tile (profiles/r2_final_ncu256_stage_stalls.txt:
			// the consumers waited for `full` in 13 % of their samples while the pipeline as a whole had throughput to spare).
			// Each warp announces its own share of the bytes (`full` counts four arrivals), so no copy can complete before its
			// expect_tx; each keeps its own copy of the halo ids (a few hundred bytes per tile) instead of sharing a ring that would
			// need one more barrier.  No warp can run a phase ahead of a slot: every fill needs every warp.
			TileMeta cur = meta_of(my_desc), nxt = cur;
			if (lane == 0) ids_fetch(cur, idring, idbar, tv.halo_cell);
			if (1 < my_tiles) nxt = meta_of(my_desc + G);
			int slot = 0;
			uint32_t use = 0;
			for (int i = 0; i < my_tiles; i++) {
				if (i + 1 < my_tiles && lane == 0) ids_fetch(nxt, idring + ((i + 1) & 1) * hpitch, idbar + ((i + 1) & 1), tv.halo_cell);
				TileMeta nn = nxt;
				if (i + 2 < my_tiles) nn = meta_of(my_desc + (size_t)(i + 2) * G);
				mbar_wait(idbar + (i & 1), (uint32_t)(i >> 1) & 1u);
				mbar_wait(empty + slot, (use & 1u) ^ 1u);
				const uint32_t q0 = smem_u32(sm + pg.off_slot + (size_t)slot * pg.slot_bytes);
				const int ngrp = (pg.dbg & 1) ? 0 : (cur.nh + 3) >> 2;
				const int mine = ngrp > pwarp ? (ngrp - pwarp + kProducerWarps - 1) / kProducerWarps : 0;   // groups pwarp, pwarp + 4, ...
				if (lane == 0) {
					const uint32_t own = (pwarp != 0 || (pg.dbg & 2)) ? 0u : (uint32_t)pg.box_cells * (QB + VB);
					mbar_arrive_expect_tx(full + slot, own + (uint32_t)mine * 4u * (QB + VB));
					if (own) {
						tma_load_2d(q0, &map_q, 0, cur.c0, full + slot);
						tma_load_2d(q0 + pg.q_bytes, &map_v, 0, cur.c0, full + slot);
					}
				}
				__syncwarp();   // the expect_tx is registered before any lane's copy can complete
				{
					const int* ids = idring + (i & 1) * hpitch;
					for (int k = lane; k < mine; k += 32) {
						const int gq = pwarp + kProducerWarps * k;
						const int4 rows = *reinterpret_cast<const int4*>(ids + 4 * gq);
						tma_gather4(q0 + (uint32_t)(pg.box_cells + 4 * gq) * QB, &gmap_q, rows, full + slot);
						tma_gather4(q0 + pg.q_bytes + (uint32_t)(pg.box_cells + 4 * gq) * VB, &gmap_v, rows, full + slot);
					}
				}
				if (PD > 0 && (i & (kProducerWarps - 1)) == pwarp && i + PD < my_tiles) prefetch_tile(my_desc + (size_t)(i + PD) * G);
				__syncwarp();   // every lane has read this tile's ids before lane 0 lets the next bulk copy overwrite the ring
				cur = nxt;
				nxt = nn;
				if (++slot == NS) {
					slot = 0;
					use++;
				}
			}
			return;
		}
		auto next_mine = [&](int i) {   // the next tile of this CTA (after i) whose slot this warp serves; >= my_tiles: none
			do i++;
			while (i < my_tiles && (i % NS) % kProducerWarps != pwarp);
			return i;
		};
		int i = next_mine(-1);
		if (i >= my_tiles) return;
		int i1 = next_mine(i), i2 = next_mine(i1);
		TileMeta cur = meta_of(my_desc + (size_t)i * G), nxt = cur;
		if (lane == 0) ids_fetch(cur, idring, idbar, tv.halo_cell);
		if (i1 < my_tiles) nxt = meta_of(my_desc + (size_t)i1 * G);
		for (int j = 0; i < my_tiles; i = i1, i1 = i2, i2 = next_mine(i2), j++) {
			const int slot = i % NS;
			const uint32_t use = (uint32_t)(i / NS);
			const bool has_next = i1 < my_tiles;
			// the next tile's halo ids start travelling, the descriptor of the tile after it is requested
			if (has_next && lane == 0) ids_fetch(nxt, idring + ((j + 1) & 1) * hpitch, idbar + ((j + 1) & 1), tv.halo_cell);
			TileMeta nn = nxt;
			if (i2 < my_tiles) nn = meta_of(my_desc + (size_t)i2 * G);
			mbar_wait(idbar + (j & 1), (uint32_t)(j >> 1) & 1u);      // this tile's halo ids are in the ring
			mbar_wait(empty + slot, (use & 1u) ^ 1u);                  // the consumers have released the slot (passes at once on its first use)
			const uint32_t q0 = smem_u32(sm + pg.off_slot + (size_t)slot * pg.slot_bytes);
			// the halo: groups of four cells per TMA gather (the plan pads every halo list to a multiple of four with cell 0)
			const int ngrp = (pg.dbg & 1) ? 0 : (cur.nh + 3) >> 2;
			if (lane == 0) {
				const uint32_t own = (pg.dbg & 2) ? 0u : (uint32_t)pg.box_cells * (QB + VB);
				mbar_arrive_expect_tx(full + slot, own + (uint32_t)ngrp * 4u * (QB + VB));
				if (own) {
					tma_load_2d(q0, &map_q, 0, cur.c0, full + slot);
					tma_load_2d(q0 + pg.q_bytes, &map_v, 0, cur.c0, full + slot);
				}
			}
			__syncwarp();   // the expect_tx is registered before any lane's copy can complete
			{
				const int* ids = idring + (j & 1) * hpitch;
				for (int gq = lane; gq < ngrp; gq += 32) {
					const int4 rows = *reinterpret_cast<const int4*>(ids + 4 * gq);
					tma_gather4(q0 + (uint32_t)(pg.box_cells + 4 * gq) * QB, &gmap_q, rows, full + slot);
					tma_gather4(q0 + pg.q_bytes + (uint32_t)(pg.box_cells + 4 * gq) * VB, &gmap_v, rows, full + slot);
				}
			}
			if (PD > 0 && i + PD < my_tiles) prefetch_tile(my_desc + (size_t)(i + PD) * G);
			__syncwarp();
			cur = nxt;
			nxt = nn;
		}
		asm volatile("cp.async.wait_all;" ::: "memory");
		return;
	}

	// ============================ consumers ====================================================================
	asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kPipeConsumerRegs));
	if ((int)blockIdx.x >= n_tiles) return;
	const int g = threadIdx.x / GT;           // consumer group
	const int tid = threadIdx.x % GT;         // thread inside the group
	const int bar_id = 1 + g;
	R* fl = reinterpret_cast<R*>(sm + pg.off_fl) + (size_t)g * NQ * pg.fmax;
	const int fmax = pg.fmax;
	int t = blockIdx.x + g * G;
	if (t >= n_tiles) return;
	// of a tile descriptor the consumers need four values
	struct Tile {
		int c0, nt, f_off, nf;
	};
	auto tile_of = [&](int tile) {
		const TileMeta mt = meta_of(tv.tiles + tile0 + tile);
		return Tile{mt.c0, tv.tiles[tile0 + tile].nt, mt.f_off, mt.nf};
	};
	Tile td = tile_of(t);
	FaceIn<R, D> cur;
	if (tid < td.nf) fetch_face<R, D, SCHEME>(tv, (size_t)td.f_off + tid, cur);
	// this group's tiles are i = g, g + 2, ... of the CTA: slot i % NS, use i / NS, kept as running values (a division per tile
	// sat in front of the wait for the slot)
	int slot = g % NS;
	uint32_t use = (uint32_t)(g / NS);
	const bool per_warp = pg.wstore || pg.direct;
	uint64_t* flfree = full + 2 * NS + g;   // (per_warp) completes once per tile of this group
	uint32_t nth = 0;                       // tiles this group has finished
	auto next_use = [&]() {
		slot += 2;
		if (slot >= NS) {
			slot -= NS;
			use++;
		}
	};
	for (;; t += 2 * G, next_use()) {
		const bool has_next = t + 2 * G < n_tiles;
		Tile tdn = td;
		if (has_next) tdn = tile_of(t + 2 * G);   // in flight during this tile
		unsigned char* Qs = sm + pg.off_slot + (size_t)slot * pg.slot_bytes;
		const unsigned char* Vs = Qs + pg.q_bytes;
		unsigned char* outb = Qs + pg.q_bytes;   // the V region is dead once phase B is over: it stages the finished Q records (un-swizzled)
		const int nf = td.nf;
		// The two groups take alternate tiles, so this group can reach a slot while its previous user (a tile of the OTHER group)
		// still waits for its data: one phase behind, which the parity of `full` alone cannot tell from "already filled again".
		// The slot's previous fill was issued before this group's last tile was, so the slot is never further behind than that:
		// wait until its previous user has released it (phase use-1 of `empty`; from then on `full` is either filling or
		// filled for THIS use, nothing else), then for this fill.
		if (use) mbar_wait(empty + slot, (use - 1u) & 1u);
		mbar_wait(full + slot, use & 1u);
		if (pg.dbg & 4) {   // timing experiment: the fill pipeline alone
			named_bar(bar_id, GT);
			if (per_warp ? (tid & 31) == 0 : tid == 0) mbar_arrive(empty + slot);
			if (!has_next) break;
			td = tdn;
			continue;
		}

		// ---- B: every face of the tile once ----------------------------------------------------------------
		if (pg.pf_face && tid + GT < nf) {
			const size_t j = (size_t)td.f_off + tid + GT;
			prefetch_l1(tv.f_idx + j);
#pragma unroll
			for (int k = 0; k < D; k++) {
				prefetch_l1(tv.fS + k * tv.T + j);
				prefetch_l1(tv.fK + k * tv.T + j);
			}
			prefetch_l1(tv.fw + j);
			prefetch_l1(tv.fdm + j);
			prefetch_l1(tv.fdi + j);
			if (SCHEME == 0) prefetch_l1(tv.fSmag + j);
		}
		for (int lf = tid; lf < nf; lf += GT) {
			// the first face of a tile was fetched a tile ago; the face tables of this tile were prefetched into the L2 by the producers
			// (a register prefetch of the next round's face measured the same and cost 19 registers and the kernel's only spills)
			if (lf != tid) fetch_face<R, D, SCHEME>(tv, (size_t)td.f_off + lf, cur);
			const int lo = (int)(cur.idx & 0xffffu), ln = (int)((cur.idx >> 16) & 0x7fffu);
			const bool ghost = (cur.idx >> 31) != 0;
			R dv[D];
#pragma unroll
			for (int k = 0; k < D; k++) dv[k] = R(0);
			if (ghost) {
				const int f = tv.f_gface[(size_t)td.f_off + lf];
#pragma unroll
				for (int k = 0; k < D; k++) dv[k] = m.d[k * m.nfs + f];
			}
			RecSide<R, D> c, n;
			c.load_q(Qs, lo);
			n.load_q(Qs, ln);
			c.load_v(Vs, lo);
			n.load_v(Vs, ln);
			R rhs[NQ];
			face_flux<R, D, SCHEME>(m.k, c, n, cur.g, ghost, dv, rhs);
			// (wstore) the flux rows still belong to the previous tile of this group until every warp has gathered from them
			if (per_warp && lf == tid && nth) mbar_wait(flfree, (nth - 1u) & 1u);
#pragma unroll
			for (int k = 0; k < NQ; k++) fl[k * fmax + lf] = rhs[k];
		}
		// (a warp without a face in this tile writes nothing: it has nothing to wait for; phases are counted by nth)

		// ---- C: ordered gather, sponge, RK update, derived values.  A cell belongs to a PAIR of lanes of one warp (l and l ^ 16):
		// the low lane takes components [0, H), the high lane [NQ - H, NQ) (the ordered sums are per component, so the split
		// changes no result; with an odd NQ the middle component is formed by both, identically) -- the SAME code for both, only
		// addresses depend on the half, so the warp does not diverge.  The lanes swap their new conservatives through the warp
		// and each writes its half of the finished record to the staging area (un-swizzled): low lane the chunks that hold
		// conservatives only, high lane -- which also forms the derived values -- the rest.  No barrier between the update and
		// the staging: the pair is one warp.
		{
			constexpr int H = (NQ + 1) / 2, EPC = Chunk16<R>::N;
			const int wl = tid & 31, part = wl >> 4, cb = part * (NQ - H);
			for (int base = 0; base < td.nt; base += GT / 2) {
				const int lc = base + (tid >> 5) * 16 + (wl & 15);
				const bool active = lc < td.nt;
				const int c = td.c0 + (active ? lc : 0);
				CellIn<R, D, H> in;
				gather_inputs<R, D, H>(m, tv, c, cb, active, first, in);
				if (base == 0) named_bar(bar_id, GT);   // every face flux of the tile is in fl (the loads above travel meanwhile)
				R mine[H];
#pragma unroll
				for (int k = 0; k < H; k++) mine[k] = R(0);
				if (active) gather_update_rec<R, D, H>(m, in, c, cb, Qs, fl, fmax, lc, dt, Ak, Bk, res, mine);
				// the record: component k is mine[k - cb] where this lane formed it, the partner's otherwise
				R rec[RC::QW];
#pragma unroll
				for (int k = 0; k < RC::QW; k++) rec[k] = R(0);
#pragma unroll
				for (int k = 0; k < H; k++) {
					const R other = pair_swap(mine[k]);
					// low lane: mine -> component k, other -> component NQ - H + k; high lane: the reverse
					rec[k] = part == 0 ? mine[k] : other;
					if (NQ - H + k >= H) rec[NQ - H + k] = part == 0 ? other : mine[k];
				}
				if (active) {
					constexpr int C0 = (NQ / EPC);   // chunks [0, C0) hold conservatives only: the low lane's share
					unsigned char* gout = reinterpret_cast<unsigned char*>(qn) + (size_t)c * QB;   // (direct) the cell's record in HBM
					if (part == 0) {
#pragma unroll
						for (int j = 0; j < C0; j++) {
							if (pg.direct)
								*reinterpret_cast<typename Chunk16<R>::T*>(gout + j * 16) = Chunk16<R>::pack(rec + j * EPC);
							else
								*reinterpret_cast<typename Chunk16<R>::T*>(outb + (size_t)lc * QB + j * 16) = Chunk16<R>::pack(rec + j * EPC);
						}
					} else {
						CellState<R, D> cs;
#pragma unroll
						for (int k = 0; k < NQ; k++) cs.q[k] = rec[k];
						derive_state<R, D, SCHEME>(m.k, cs);
						rec[RC::RHO_INV] = cs.rho_inv;
						rec[RC::RPSI] = cs.Rpsi;
						rec[RC::AUX] = cs.aux;
#pragma unroll
						for (int j = C0; j < CQ; j++) {
							if (pg.direct)
								*reinterpret_cast<typename Chunk16<R>::T*>(gout + j * 16) = Chunk16<R>::pack(rec + j * EPC);
							else
								*reinterpret_cast<typename Chunk16<R>::T*>(outb + (size_t)lc * QB + j * 16) = Chunk16<R>::pack(rec + j * EPC);
						}
					}
				}
			}
		}
		// the first face of this group's next tile travels during the store and the wait for its slot
		if (has_next && tid < tdn.nf) fetch_face<R, D, SCHEME>(tv, (size_t)tdn.f_off + tid, cur);
		if (pg.direct) {
			// nothing is staged: a warp is done with the slot and the flux rows when its lanes are
			__syncwarp();
			if ((tid & 31) == 0) {
				mbar_arrive(empty + slot);
				mbar_arrive(flfree);
			}
			nth++;
			if (!has_next) break;
			td = tdn;
			continue;
		}
		fence_proxy_async();   // the staging writes above become visible to the TMA engine
		if (pg.wstore) {
			// every warp hands over the records of its own cells (16 per pass, contiguous in the staging area and in HBM) and
			// releases the slot and the flux rows for itself: nobody waits for the slowest warp of the group here
			__syncwarp();
			if ((tid & 31) == 0) {
				for (int base = (tid >> 5) * 16; base < td.nt; base += GT / 2) {
					const int n = min(16, td.nt - base);
					bulk_store(reinterpret_cast<unsigned char*>(qn) + (size_t)(td.c0 + base) * QB, smem_u32(outb) + (uint32_t)base * QB, (uint32_t)n * QB);
				}
				bulk_commit();
				bulk_wait_read0();
				mbar_arrive(empty + slot);
				mbar_arrive(flfree);
			}
			__syncwarp();
			nth++;
		} else {
			named_bar(bar_id, GT);
			if (tid == 0) {
				bulk_store(reinterpret_cast<unsigned char*>(qn) + (size_t)td.c0 * QB, smem_u32(outb), (uint32_t)td.nt * QB);
				bulk_commit();
				bulk_wait_read0();           // the store has read its source: the slot may be refilled
				mbar_arrive(empty + slot);   // every thread of the group is past its last access to the slot
			}
		}
		if (!has_next) break;
		td = tdn;
	}
	if (!pg.direct && (pg.wstore ? (tid & 31) == 0 : tid == 0)) bulk_wait0();   // the last store has left shared memory and is performed before the CTA exits
}

// ---------------------------------------------------------------------------------------------------------------------
// calc_VIS (cfd_v0.cpp:1744-1860) as a persistent, TMA-fed kernel: same producer / ring as k_stage_pipe, with
//   * per slot: the Q records of tile + halo (TMA box + 16-byte cp.async, swizzled) and the tile's slice of the tile-ordered
//     face tables S[D], w, idx (contiguous per tile: D + 2 linear bulk copies, cp.async.bulk.shared::cta.global);
//   * three consumer groups of 128 threads on alternate tiles: one thread per cell walks the cell's faces in ascending face
//     id (the reference's summation order), forms dudx / dTdx / sigmaU / tr and writes the cell's V record into the slot's
//     (by then dead) Q region, from where ONE bulk store takes the tile's V records to HBM.
// ---------------------------------------------------------------------------------------------------------------------
struct GradGeom {
	int n_slots, box_cells, smax, fmax, hmax;
	uint32_t off_bar, off_meta, off_ids, off_slot;
	uint32_t slot_bytes;    // Q region (also the V staging of the finished tile) + face region
	uint32_t q_bytes;       // Q region: max(smax * QB, box_cells * VB) rounded up to 1024
	int pf_dist;            // L2 prefetch distance (see PipeGeom)
	int dbg;                // timing experiments (see PipeGeom)
	int wstore;             // per-warp stores and releases (see PipeGeom)
	int direct;             // 1: every thread writes its cell's V record from registers straight to HBM (eight 16-byte stores; the L2 merges the
	                        // half sectors): no staging -- whose 128-byte-strided writes were 8-way bank conflicts, more than half of the
	                        // kernel's shared-memory wavefronts (profiles/r2_final_ncu256_summary.txt) --, no proxy fence, no group barrier
};
// Ring depth and group count are tied: a group reaches use k of a slot knowing only that ITS OWN earlier tiles were
// released; the parity wait for the release of use k - 1 is unambiguous only if use k - 2 (tile i - 2 NS) was one of them,
// i.e. 2 NS must be a multiple of the number of groups (k_stage_pipe: two groups, any depth; here three groups: depth 3, 6, 9;
// four groups: an even depth).
// Group count NG: 3 (512 threads: every thread may use 128 registers, no setmaxnreg) or 4 (640 threads on the stage kernel's register
// budget: 96 at launch, the producers give back all but 32, the consumers grow to 112 -- a third more warps for a kernel whose
// consumers wait on their own dependent latencies).
constexpr int kGradGroupThreads = 128;
__host__ __device__ constexpr int grad_threads_total(int ng) { return ng * kGradGroupThreads + kPipeProducerThreads; }

__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <class R, int D, int NG>
__global__ void __launch_bounds__(grad_threads_total(NG), 1)
    k_grad_pipe(DevMesh<R> m, TileView<R> tv, const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap gmap_q, GradGeom pg, const R* __restrict__ q, int tile0,
                int n_tiles) {
	using RC = Rec<D>;
	constexpr int QB = RC::QW * (int)sizeof(R), VB = RC::VW * (int)sizeof(R), CVC = VB / 16;
	constexpr int GT = kGradGroupThreads;
	static_assert(NG == 3 || (NG == 4 && kPipeProducerThreads * kPipeProducerRegs + NG * GT * kPipeConsumerRegs <= grad_threads_total(NG) * 96), "register budget of the four-group variant");
	extern __shared__ unsigned char smem_dyn[];
	unsigned char* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
	uint64_t* full = reinterpret_cast<uint64_t*>(sm + pg.off_bar);
	uint64_t* empty = full + pg.n_slots;
	const int NS = pg.n_slots, G = gridDim.x;
	const int lane = threadIdx.x & 31;
	const uint32_t frow = (uint32_t)pg.fmax * (uint32_t)sizeof(R);   // one face row: S_k or w

	if (threadIdx.x == 0) {
		for (int s = 0; s < NS; s++) {
			mbar_init(full + s, 1);
			mbar_init(empty + s, (pg.wstore || pg.direct) ? GT / 32 : 1);
		}
		for (int k = 0; k < 2 * kProducerWarps; k++) mbar_init(reinterpret_cast<uint64_t*>(sm + pg.off_meta) + k, 1);   // the producers' id rings
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	if ((int)blockIdx.x >= n_tiles) return;   // (the whole CTA)

	if (threadIdx.x >= NG * GT) {
		// ============================ producers (roles as in k_stage_pipe) ========================================
		if constexpr (NG == 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kPipeProducerRegs));
		const int pwarp = (threadIdx.x - NG * GT) >> 5;
		const int hpitch = (pg.hmax + 3) & ~3;
		int* idring = reinterpret_cast<int*>(sm + pg.off_ids) + pwarp * 2 * hpitch;
		uint64_t* idbar = reinterpret_cast<uint64_t*>(sm + pg.off_meta) + pwarp * 2;
		const unsigned char* qsrc = reinterpret_cast<const unsigned char*>(q);
		const int my_tiles = (n_tiles - (int)blockIdx.x + G - 1) / G;
		const TileDesc* my_desc = tv.tiles + tile0 + blockIdx.x;
		auto next_mine = [&](int i) {   // the next tile of this CTA (after i) whose slot this warp serves; >= my_tiles: none
			do i++;
			while (i < my_tiles && (i % NS) % kProducerWarps != pwarp);
			return i;
		};
		int i = next_mine(-1);
		if (i >= my_tiles) return;
		int i1 = next_mine(i), i2 = next_mine(i1);
		TileMeta cur = meta_of(my_desc + (size_t)i * G), nxt = cur;
		if (lane == 0) ids_fetch(cur, idring, idbar, tv.halo_cell);
		if (i1 < my_tiles) nxt = meta_of(my_desc + (size_t)i1 * G);
		const int PD = pg.pf_dist;
		for (int j = 0; i < my_tiles; i = i1, i1 = i2, i2 = next_mine(i2), j++) {
			const int slot = i % NS;
			const uint32_t use = (uint32_t)(i / NS);
			const bool has_next = i1 < my_tiles;
			if (has_next && lane == 0) ids_fetch(nxt, idring + ((j + 1) & 1) * hpitch, idbar + ((j + 1) & 1), tv.halo_cell);
			TileMeta nn = nxt;
			if (i2 < my_tiles) nn = meta_of(my_desc + (size_t)i2 * G);
			mbar_wait(idbar + (j & 1), (uint32_t)(j >> 1) & 1u);
			mbar_wait(empty + slot, (use & 1u) ^ 1u);
			const uint32_t q0 = smem_u32(sm + pg.off_slot + (size_t)slot * pg.slot_bytes);
			const uint32_t fs = q0 + pg.q_bytes;
			// the tile's slice of the face tables starts on a 16-byte boundary (the plan pads f_off to a multiple of 4) and is
			// copied in whole 16-byte units (the tables are padded behind the last tile)
			const uint32_t nf4 = (uint32_t)(cur.nf + 3) & ~3u;
			const int ngrp = (pg.dbg & 1) ? 0 : (cur.nh + 3) >> 2;
			if (lane == 0) {
				mbar_arrive_expect_tx(full + slot, ((pg.dbg & 2) ? 0u : (uint32_t)pg.box_cells * QB) + nf4 * (uint32_t)((D + 1) * sizeof(R) + 4) + (uint32_t)ngrp * 4u * QB);
				if (!(pg.dbg & 2)) tma_load_2d(q0, &map_q, 0, cur.c0, full + slot);
			}
			__syncwarp();   // the expect_tx precedes the bulk copies of the other lanes
			if (nf4) {
				if (lane >= 1 && lane < 1 + D) bulk_load(fs + (uint32_t)(lane - 1) * frow, tv.fS + (size_t)(lane - 1) * tv.T + cur.f_off, nf4 * (uint32_t)sizeof(R), full + slot);
				if (lane == 1 + D) bulk_load(fs + (uint32_t)D * frow, tv.fw + cur.f_off, nf4 * (uint32_t)sizeof(R), full + slot);
				if (lane == 2 + D) bulk_load(fs + (uint32_t)(D + 1) * frow, tv.f_idx + cur.f_off, nf4 * 4u, full + slot);
			}
			{
				const int* ids = idring + (j & 1) * hpitch;
				for (int gq = lane; gq < ngrp; gq += 32) tma_gather4(q0 + (uint32_t)(pg.box_cells + 4 * gq) * QB, &gmap_q, *reinterpret_cast<const int4*>(ids + 4 * gq), full + slot);
			}
			if (PD > 0 && lane >= 8 && lane < 16 && i + PD < my_tiles) {
				const TileMeta pm = meta_of(my_desc + (size_t)(i + PD) * G);
				const int ac0 = pm.c0, af = pm.f_off;
				const uint32_t pf4 = (uint32_t)(pm.nf + 3) & ~3u;
				if (lane == 8) bulk_prefetch_l2(qsrc + (size_t)ac0 * QB, (uint32_t)pg.box_cells * QB);
				if (pf4) {
					if (lane == 9) bulk_prefetch_l2(tv.f_idx + af, pf4 * 4u);
					if (lane >= 10 && lane < 10 + D) bulk_prefetch_l2(tv.fS + (size_t)(lane - 10) * tv.T + af, pf4 * (uint32_t)sizeof(R));
					if (lane == 10 + D) bulk_prefetch_l2(tv.fw + af, pf4 * (uint32_t)sizeof(R));
				}
			}
			cur = nxt;
			nxt = nn;
		}
		asm volatile("cp.async.wait_all;" ::: "memory");
		return;
	}

	// ============================ consumers ====================================================================
	if constexpr (NG == 4) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kPipeConsumerRegs));
	const int g = threadIdx.x / GT, tid = threadIdx.x % GT;
	const int bar_id = 1 + g;
	int t = blockIdx.x + g * G;
	if (t >= n_tiles) return;
	// what the gather needs from global memory travels one tile ahead: this tile's gather lists and cell volume were requested
	// while the previous tile was computed, the next tile's descriptor is requested now
	auto cell_inputs = [&](const TileDesc& d, int* e, R& vinv) {
		const bool act = tid < d.nt;
		const int c = d.c0 + (act ? tid : 0);
#pragma unroll
		for (int s = 0; s < kMaxSlots; s++) e[s] = (act && s < m.F) ? (int)tv.csr_local[(size_t)s * m.n_cells + c] : 0;
		vinv = act ? m.vol_inv[c] : R(0);
	};
	TileDesc td = tv.tiles[tile0 + t];
	TileDesc tdn = td;
	if (t + NG * G < n_tiles) tdn = tv.tiles[tile0 + t + NG * G];
	int e[kMaxSlots];
	R vinv;
	cell_inputs(td, e, vinv);
	int slot = g % NS;   // this group's tiles are i = g, g + NG, ... of the CTA: slot i % NS, use i / NS as running values (NG <= NS)
	uint32_t use = (uint32_t)(g / NS);
	auto next_use = [&]() {
		slot += NG;
		if (slot >= NS) {
			slot -= NS;
			use++;
		}
	};
	for (;; t += NG * G, next_use()) {
		const bool has_next = t + NG * G < n_tiles;
		TileDesc tdnn = tdn;
		if (t + 2 * NG * G < n_tiles) tdnn = tv.tiles[tile0 + t + 2 * NG * G];
		unsigned char* Qs = sm + pg.off_slot + (size_t)slot * pg.slot_bytes;
		const unsigned char* Fs = Qs + pg.q_bytes;
		const R* fg = reinterpret_cast<const R*>(Fs);                                         // [D+1][fmax]: S, w
		const uint32_t* fi = reinterpret_cast<const uint32_t*>(Fs + (size_t)(D + 1) * frow);  // [fmax]
		const int fmax = pg.fmax;
		if (use) mbar_wait(empty + slot, (use - 1u) & 1u);   // see k_stage_pipe: the groups run up to NG - 1 tiles apart (NS >= NG)
		// one thread per cell (the host launches this kernel for tiles of at most GT cells)
		const int lc = tid;
		const bool active = lc < td.nt;
		mbar_wait(full + slot, use & 1u);
		if (pg.dbg & 4) {   // timing experiment: the fill pipeline alone
			named_bar(bar_id, GT);
			if ((pg.wstore || pg.direct) ? (tid & 31) == 0 : tid == 0) mbar_arrive(empty + slot);
			if (!has_next) break;
			td = tdn;
			tdn = tdnn;
			continue;
		}
		R rec[RC::VW];
		if (active) {
			RecSide<R, D> own;
			own.load_q(Qs, lc);
			R cU[D];
#pragma unroll
			for (int k = 0; k < D; k++) cU[k] = own.q(k + 1) * own.rho_inv();
			const R c_Rpsi = own.Rpsi();
			R dudx[D][D], dTdx[D], sigmaU[D];
#pragma unroll
			for (int a = 0; a < D; a++) {
				dTdx[a] = R(0);
#pragma unroll
				for (int b = 0; b < D; b++) dudx[a][b] = R(0);
			}
#pragma unroll
			for (int s = 0; s < kMaxSlots; s++) {
				if (e[s] != 0) {   // (dense list; no early exit: the loads of all six faces can travel together)
				const bool is_own = e[s] > 0;
				const int lf = (is_own ? e[s] : -e[s]) - 1;
				const uint32_t idx = fi[lf];
				const int lo = is_own ? (int)((idx >> 16) & 0x7fffu) : (int)(idx & 0xffffu);   // the other side
				RecSide<R, D> oth;
				oth.load_q(Qs, lo);
				R oU[D];
#pragma unroll
				for (int k = 0; k < D; k++) oU[k] = oth.q(k + 1) * oth.rho_inv();
				const R o_Rpsi = oth.Rpsi();
				const R w = fg[D * fmax + lf];
				R face_U[D], face_T, sov[D];
				if (is_own) {
					grad_face_values<R, D>(m.k, w, cU, c_Rpsi, oU, o_Rpsi, face_U, face_T);
#pragma unroll
					for (int a = 0; a < D; a++) sov[a] = fg[a * fmax + lf] * vinv;
				} else {
					grad_face_values<R, D>(m.k, w, oU, o_Rpsi, cU, c_Rpsi, face_U, face_T);
#pragma unroll
					for (int a = 0; a < D; a++) sov[a] = -fg[a * fmax + lf] * vinv;
				}
#pragma unroll
				for (int a = 0; a < D; a++) {
#pragma unroll
					for (int b = 0; b < D; b++) dudx[a][b] += face_U[a] * sov[b];
					dTdx[a] += face_T * sov[a];
				}
				}
			}
			// the tau / sigmaU block of calc_VIS (U = rhoU / rho: a true division, cfd_v0.cpp:1806-1857)
			R Ud[D], tau[D][D];
#pragma unroll
			for (int a = 0; a < D; a++) Ud[a] = own.q(a + 1) / own.q(0);
			stress<R, D>(m.k, dudx, tau);
#pragma unroll
			for (int a = 0; a < D; a++) sigmaU[a] = dotD<R, D>(Ud, tau[a]);
#pragma unroll
			for (int k = 0; k < RC::VW; k++) rec[k] = R(0);
#pragma unroll
			for (int a = 0; a < D; a++) {
#pragma unroll
				for (int b = 0; b < D; b++) rec[RC::DUDX + a * D + b] = dudx[a][b];
				rec[RC::DTDX + a] = dTdx[a];
				rec[RC::SIGMAU + a] = sigmaU[a];
			}
			if (D == 3) rec[RC::TR] = dudx_trace_neg<R, D>(&dudx[0][0]);
		}
		if (has_next) cell_inputs(tdn, e, vinv);   // the next tile's gather lists and volumes travel from here on
		if (pg.direct) {
			if (active) {
				unsigned char* gout = reinterpret_cast<unsigned char*>(m.vis) + (size_t)(td.c0 + lc) * VB;
#pragma unroll
				for (int j = 0; j < CVC; j++) *reinterpret_cast<typename Chunk16<R>::T*>(gout + j * 16) = Chunk16<R>::pack(rec + j * Chunk16<R>::N);
			}
			__syncwarp();   // every lane of the warp has read what it needs of the slot
			if ((tid & 31) == 0) mbar_arrive(empty + slot);
			if (!has_next) break;
			td = tdn;
			tdn = tdnn;
			continue;
		}
		named_bar(bar_id, GT);   // every thread of the group has read what it needs of the Q region: it becomes the V staging
		if (active) {
#pragma unroll
			for (int j = 0; j < CVC; j++) *reinterpret_cast<typename Chunk16<R>::T*>(Qs + (size_t)lc * VB + j * 16) = Chunk16<R>::pack(rec + j * Chunk16<R>::N);
		}
		fence_proxy_async();
		if (pg.wstore) {
			// every warp hands over the V records of its own 32 cells and releases the slot for itself (`empty` counts the warps of
			// the group): no second group barrier, no single thread the whole group waits for
			__syncwarp();
			if ((tid & 31) == 0) {
				const int base = tid, n = min(32, td.nt - base);
				if (n > 0) {
					bulk_store(reinterpret_cast<unsigned char*>(m.vis) + (size_t)(td.c0 + base) * VB, smem_u32(Qs) + (uint32_t)base * VB, (uint32_t)n * VB);
					bulk_commit();
					bulk_wait_read0();
				}
				mbar_arrive(empty + slot);
			}
			__syncwarp();
		} else {
			named_bar(bar_id, GT);
			if (tid == 0) {
				bulk_store(reinterpret_cast<unsigned char*>(m.vis) + (size_t)td.c0 * VB, smem_u32(Qs), (uint32_t)td.nt * VB);
				bulk_commit();
				bulk_wait_read0();   // the slot may be refilled once the store has read it
				mbar_arrive(empty + slot);
			}
		}
		if (!has_next) break;
		td = tdn;
		tdn = tdnn;
	}
	if (!pg.direct && (pg.wstore ? (tid & 31) == 0 : tid == 0)) bulk_wait0();
}

}  // namespace lfm
