// Appearance head on the 5th-generation tensor cores, second design: TWO MLP groups per SM with every layer's A operand in
// TENSOR MEMORY (sm_100a only).  Same arithmetic as k_app_tc (tvm_mlp_tc.cu): tensoRF.py:228-244 (gather + basis_mat),
// tensorBase.py:9-15,62-86 (positional encoding + MLPRender_Fea), REFTensoRF.py:5-29,107-133,216-238 (variant).
//
// Why: k_app_tc is one dependent chain GEMM0 -> epi0 -> GEMM1 -> epi1 -> GEMM2 -> epi2 -> GEMM3 per 128-entry tile (~7200
// cycles, tensor pipe 18 % busy) next to a gather group that shares the L1 data pipe with the chain's own shared-memory
// traffic (1700 activation-store wavefronts + 1900 operand-fetch wavefronts per tile) and is left with 17 KB of L1 by the
// 211 KB of shared memory (profiles/r01q_notes.txt).  Here
//   * the activations of layers 1..3 never touch shared memory: a thread (= tile row = TMEM lane) reads its accumulator row
//     with tcgen05.ld, applies PE / ReLU, and writes the packed 16-bit row back with tcgen05.st; the next GEMM is
//     tcgen05.mma [d], [a_tmem], b_desc.  Shared memory holds the weight image (86 KB) and two GEMM0 stages (72 KB) only,
//     so ~70 KB stay L1 for the gather, and the A-side operand fetches + activation stores leave the L1 data pipe;
//   * the freed 40 KB and the 512 TMEM columns carry a second, independent MLP group: while group 0 sits in an epilogue,
//     group 1's MMAs run, and vice versa (tiles alternate between the groups; stage s of the GEMM0 operand belongs to
//     group s);
//   * GEMM0 of a group's NEXT tile is issued as soon as its epilogue 0 has drained the basis accumulator, i.e. two layers
//     ahead of its use.
// TMEM columns of group g (base 256 g): [0,128) layer accumulator D (GEMM3's 16 output columns reuse [0,16)),
// [128,208) A operand (K <= 160 16-bit elements), [208,256) basis / stacked-head accumulator (NH <= 48).
#include "tvm_tc.cuh"

namespace tvm {

using namespace tc;

namespace app2 {

constexpr int kGroups = 2;
constexpr int kMlpWarps2 = 4 * kGroups;
constexpr uint32_t kColD = 0, kColA = 128, kColBas = 208, kGroupCols = 256;
// K-chunk stride of the GEMM0 operand image: 128 rows x 16 B.  (Padding it by 64 B removes the two-way bank conflict of the
// gather's 8-byte stores -- 5 instead of 2.25 wavefronts per entry -- but the tensor core then reads core matrices that
// straddle 128-byte lines: 1.85 instead of 1.48 ms per frame, profiles/r02_notes.txt N.)
constexpr uint32_t kLboA0 = kRows * 16;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }

// One tile row of the GEMM0 operand: 48-channel plane x line products of entry e (4 lanes per entry, lane q owns channels
// 4q..4q+3 of every 16), written as 16-bit pairs into the K-major core-matrix image (row stride 16 B, K-chunk stride LBO).
template <int CA, bool PB16, bool H16>
__device__ __forceinline__ void gather_row(const FwdParams& P, uint8_t* arow, bool live, const float4 uw, int q) {
  constexpr uint32_t LBO_A = kLboA0;
  const TvmModel& m = P.m;
  if (live) {
    const float u[3] = {uw.x, uw.y, uw.z};
    if (PB16) {
      AxisPair ax[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) ax[i] = axis_pair(u[i], m.grid[i]);
#pragma unroll
      for (int kk = 0; kk < 3; ++kk) {
        const VmPair t = vm_pair(m, ax, kk, CA);
        // pair records: [texel w c..c+3 | texel w+1 c..c+3] per 16 bytes, so one load brings both taps of the pair
        const uint4* pl = reinterpret_cast<const uint4*>(m.app_plane_pair[kk]);
        const uint4* ln = reinterpret_cast<const uint4*>(m.app_line_pair[kk]);
        // every texel unpacked to fp32 (HADD2.F32), fp32 weights: 8 + 4 + 1 instructions per channel (gather_row_mixed: 4 + 2 + 1)
#pragma unroll
        for (int c = q * 4; c < CA; c += 16) {
          const uint4 r0 = __ldg(pl + ((t.row0 + c) >> 2));
          const uint4 r1 = __ldg(pl + ((t.row1 + c) >> 2));
          const uint4 lr = __ldg(ln + ((t.lrow + c) >> 2));
          const float2 a0 = unpack16<H16>(r0.x), a1 = unpack16<H16>(r0.y), b0 = unpack16<H16>(r0.z), b1 = unpack16<H16>(r0.w);
          const float2 c0 = unpack16<H16>(r1.x), c1 = unpack16<H16>(r1.y), d0 = unpack16<H16>(r1.z), d1 = unpack16<H16>(r1.w);
          const float2 l00 = unpack16<H16>(lr.x), l01 = unpack16<H16>(lr.y), l10 = unpack16<H16>(lr.z), l11 = unpack16<H16>(lr.w);
          const float px = a0.x * t.nw + b0.x * t.ne + c0.x * t.sw + d0.x * t.se;
          const float py = a0.y * t.nw + b0.y * t.ne + c0.y * t.sw + d0.y * t.se;
          const float pz = a1.x * t.nw + b1.x * t.ne + c1.x * t.sw + d1.x * t.se;
          const float pw = a1.y * t.nw + b1.y * t.ne + c1.y * t.sw + d1.y * t.se;
          const float lx = l00.x * t.lw0 + l10.x * t.lw1, ly = l00.y * t.lw0 + l10.y * t.lw1;
          const float lz = l01.x * t.lw0 + l11.x * t.lw1, lw = l01.y * t.lw0 + l11.y * t.lw1;
          const int k = kk * CA + c;
          const uint2 packed = make_uint2(pack16<H16>(px * lx, py * ly), pack16<H16>(pz * lz, pw * lw));
          *reinterpret_cast<uint2*>(arow + (k >> 3) * LBO_A + (k & 7) * 2) = packed;
        }
      }
    } else {
      Axis ax[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) ax[i] = axis_taps(u[i], m.grid[i]);
#pragma unroll
      for (int kk = 0; kk < 3; ++kk) {
        const VmTaps t = vm_taps(m, ax, kk);
#pragma unroll
        for (int c = q * 4; c < CA; c += 16) {
          float4 pv, lv;
          vm_sample4(m.app_plane[kk], m.app_line[kk], t, CA, c, pv, lv);
          const int k = kk * CA + c;
          const uint2 packed = make_uint2(pack16<H16>(pv.x * lv.x, pv.y * lv.y), pack16<H16>(pv.z * lv.z, pv.w * lv.w));
          *reinterpret_cast<uint2*>(arow + (k >> 3) * LBO_A + (k & 7) * 2) = packed;
        }
      }
    }
  } else {
#pragma unroll
    for (int kk = 0; kk < 3; ++kk)
#pragma unroll
      for (int c = q * 4; c < CA; c += 16) {
        const int k = kk * CA + c;
        *reinterpret_cast<uint2*>(arow + (k >> 3) * LBO_A + (k & 7) * 2) = make_uint2(0u, 0u);
      }
  }
}

// The same tile row from the 16-bit pair records with MIXED-PRECISION FMAs and the per-entry setup shared by the four lanes.
//  * fma.rn.f32.f16 / .bf16 (sm_100, SASS FHFMA): fp32 += 16 bit x 16 bit, each factor either half of a 32-bit register.  The
//    gathered texels are multiplied as they were loaded, without an unpack instruction each: 4 + 2 + 1 instructions per channel
//    instead of 8 + 4 + 1.  The instruction takes BOTH factors in 16 bits, so the six interpolation weights of a pair are put on
//    the 2^-Q lattice (Q = 11 for fp16, 8 for bf16: every lattice point of [0, 1] is exact in the format) by telescoped
//    rounding, which keeps their sums: the four plane weights still add up to the rounded total, the two of a texel column to
//    the rounded column weight.  Texel x weight is then exact in fp32; what changes against fp32 weights is |dw| <= 2^-Q per tap
//    with sum(dw) = 0 -- a shift of the sample by at most 2^-Q texel, an error proportional to the DIFFERENCE of neighbouring
//    texels, below the rounding of the 16-bit texels themselves.
//  * lane q < 3 of an entry prepares pair q only (axes, row offsets, lattice weights: what all four lanes used to repeat for
//    all three pairs) and hands 5 registers to its neighbours by shuffle; the 27 loads of a lane take their channel offsets as
//    immediates from 9 base addresses.
// Gather loop: 755 -> ~480 SASS instructions per pass of 8 entries.
template <int CA, bool H16>
__device__ __forceinline__ void gather_row_mixed(const FwdParams& P, uint8_t* arow, bool live, const float4 uw, int lane) {
  constexpr uint32_t LBO_A = kLboA0;
  constexpr uint32_t C4 = CA / 4;                 // 16-byte groups per texel record
  const TvmModel& m = P.m;
  const int q = lane & 3, kp = min(q, 2);         // this lane prepares pair kp (lane 3 repeats pair 2; nobody reads its copy)
  uint32_t row0, lrow, w_a, w_b, w_c;
  {
    const float u_w = kp == 2 ? uw.y : uw.x, u_h = kp == 0 ? uw.y : uw.z, u_l = kp == 0 ? uw.z : (kp == 1 ? uw.y : uw.x);
    const int g_w = kp == 2 ? m.grid[1] : m.grid[0], g_h = kp == 0 ? m.grid[1] : m.grid[2];
    const int g_l = kp == 0 ? m.grid[2] : (kp == 1 ? m.grid[1] : m.grid[0]);
    const AxisPair aw = axis_pair(u_w, g_w), ah = axis_pair(u_h, g_h), al = axis_pair(u_l, g_l);
    row0 = ((uint32_t)ah.b * (uint32_t)g_w + (uint32_t)aw.b) * C4;
    lrow = (uint32_t)al.b * C4;
    constexpr float kQ = H16 ? 2048.0f : 256.0f, kInvQ = 1.0f / kQ;
    const float nw = TVM_MUL(aw.p0, ah.p0), ne = TVM_MUL(aw.p1, ah.p0), sw = TVM_MUL(aw.p0, ah.p1), se = TVM_MUL(aw.p1, ah.p1);
    const float q_se = rintf(se * kQ), q_e = rintf((ne + se) * kQ), q_s = rintf((sw + se) * kQ);
    const float q_all = rintf(((nw + ne) + (sw + se)) * kQ);
    const float q_ne = q_e - q_se, q_sw = q_s - q_se, q_nw = q_all - q_e - q_sw;
    const float q_l1 = rintf(al.p1 * kQ), q_l0 = rintf((al.p0 + al.p1) * kQ) - q_l1;
    w_a = pack16<H16>(q_nw * kInvQ, q_ne * kInvQ);      // lattice values: exact in the 16-bit format
    w_b = pack16<H16>(q_sw * kInvQ, q_se * kInvQ);
    w_c = pack16<H16>(q_l0 * kInvQ, q_l1 * kInvQ);
    if (!live) w_a = w_b = w_c = 0u;                    // rows behind the last entry: zeros (their loads hit texel 0)
  }
#pragma unroll
  for (int kk = 0; kk < 3; ++kk) {
    const int src = (lane & 28) | kk;
    const uint32_t r0 = __shfl_sync(0xffffffffu, row0, src), lr = __shfl_sync(0xffffffffu, lrow, src);
    const uint32_t wa = __shfl_sync(0xffffffffu, w_a, src), wb = __shfl_sync(0xffffffffu, w_b, src);
    const uint32_t wc = __shfl_sync(0xffffffffu, w_c, src);
    // pair records: [texel w c..c+3 | texel w+1 c..c+3] per 16 bytes; lane q owns group q of every four
    const uint4* p0 = reinterpret_cast<const uint4*>(m.app_plane_pair[kk]) + r0 + q;
    const uint4* p1 = p0 + (uint32_t)m.grid[mat0(kk)] * C4;
    const uint4* pl = reinterpret_cast<const uint4*>(m.app_line_pair[kk]) + lr + q;
    const uint16_t w_nw = half_of(wa, false), w_ne = half_of(wa, true), w_sw = half_of(wb, false), w_se = half_of(wb, true);
    const uint16_t w_l0 = half_of(wc, false), w_l1 = half_of(wc, true);
#pragma unroll
    for (int i = 0; i < CA / 16; ++i) {
      const uint4 t0 = __ldg(p0 + 4 * i), t1 = __ldg(p1 + 4 * i), tl = __ldg(pl + 4 * i);
      // t.x = texel w, channels c, c+1; t.y = texel w, channels c+2, c+3; t.z / t.w = the same of texel w+1
      auto plane = [&](uint32_t nw2, uint32_t ne2, uint32_t sw2, uint32_t se2, bool hi) {
        return fhfma<H16>(half_of(nw2, hi), w_nw, fhfma<H16>(half_of(ne2, hi), w_ne,
               fhfma<H16>(half_of(sw2, hi), w_sw, fhfma<H16>(half_of(se2, hi), w_se, 0.0f))));
      };
      auto line = [&](uint32_t l02, uint32_t l12, bool hi) {
        return fhfma<H16>(half_of(l02, hi), w_l0, fhfma<H16>(half_of(l12, hi), w_l1, 0.0f));
      };
      const float vx = plane(t0.x, t0.z, t1.x, t1.z, false) * line(tl.x, tl.z, false);
      const float vy = plane(t0.x, t0.z, t1.x, t1.z, true) * line(tl.x, tl.z, true);
      const float vz = plane(t0.y, t0.w, t1.y, t1.w, false) * line(tl.y, tl.w, false);
      const float vw = plane(t0.y, t0.w, t1.y, t1.w, true) * line(tl.y, tl.w, true);
      const int k = kk * CA + q * 4 + 16 * i;
      *reinterpret_cast<uint2*>(arow + (k >> 3) * LBO_A + (k & 7) * 2) = make_uint2(pack16<H16>(vx, vy), pack16<H16>(vz, vw));
    }
  }
}

// NGW gather warps next to the 8 MLP warps.  Registers: the launch allocates L = floor8(65536 / threads) per thread; then
// the gather warps release and the MLP warps take registers with setmaxnreg.  The MLP warps can only take what the gather
// warps released (the CTA pool; the SM's unallocated remainder is NOT in it -- a larger request deadlocks):
//   8 * (kMlp - L) <= NGW * (L - kGather).   The MLP code needs 112 registers (no spills, checked in the SASS).
template <int NGW> struct RegSplit;
#if defined(TVM_APP2_MLP_REGS) && defined(TVM_APP2_GATHER_REGS)
template <int NGW> struct RegSplit { static constexpr int kMlp = TVM_APP2_MLP_REGS, kGather = TVM_APP2_GATHER_REGS; };
#else
template <> struct RegSplit<8> { static constexpr int kMlp = 152, kGather = 104; };      // L = 128: 8*24 = 8*24
template <> struct RegSplit<12> { static constexpr int kMlp = 120, kGather = 80; };      // L =  96: 8*24 = 12*16
template <> struct RegSplit<16> { static constexpr int kMlp = 112, kGather = 64; };      // L =  80: 8*32 = 16*16
#endif

// EV: TVM_EVAL_ONLY launch (compositing inside the head; a template parameter so that the stash instantiation keeps its registers)
template <int CA, int APP_DIM, int FEA_PE, int VIEW_PE, bool REF, bool PB16, bool H16, int NGW, bool EV>
__global__ void __launch_bounds__((kMlpWarps2 + NGW) * 32, 1) k_app_tc2(const FwdParams P) {
  static_assert(!(EV && REF), "TVM_EVAL_ONLY is not offered for TVM_VARIANT_REF");
  constexpr int kThreads = (kMlpWarps2 + NGW) * 32;
  constexpr int NH = REF ? TVM_REF_HEAD_LD : 32;
  constexpr int C0 = REF ? 1 : 0;                       // REF: column 0 of the MLP input is -dot (REFTensoRF.py:20)
  constexpr int IN_C = 2 * VIEW_PE * 3 + 2 * FEA_PE * APP_DIM + 3 + APP_DIM + C0;
  constexpr int K0 = 3 * CA;
  constexpr int K1 = (IN_C + 15) / 16 * 16;
  static_assert(FEA_PE == 2 && VIEW_PE == 2, "the register-resident PE builder is written for 2 frequencies");
  static_assert(K0 % 16 == 0 && APP_DIM <= 32 && (!REF || APP_DIM + 8 <= NH), "unsupported shape");
  static_assert(IN_C < K1 && K1 <= 160 && K1 % 32 == 0, "A operand: 80 TMEM columns, written 16 at a time; constant-one column in the K padding");
  static_assert(NH <= 48, "basis accumulator: 48 TMEM columns");

  extern __shared__ __align__(1024) uint8_t smem[];
  const Image img(CA, IN_C, NH);
  uint8_t* sW = smem;                                          // weight image (16-bit operands + fp32 tail)
  uint8_t* sA0 = smem + ((img.bytes_fwd + 1023) & ~1023u);     // 2 stages of the GEMM0 operand [128 x K0]; stage s feeds group s
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA0 + 2 * (K0 / 8) * kLboA0);
  uint64_t* full = bars;            // [2] stage filled by the gather warps
  uint64_t* empty = bars + 2;       // [2] stage consumed by GEMM0 (tcgen05.commit)
  uint64_t* mma_bars = bars + 4;    // [2] layer GEMM of group g complete
  uint64_t* bas_bars = bars + 6;    // [2] GEMM0 of group g complete
  uint64_t* w_bar = bars + 8;       // weight image landed (cp.async.bulk transaction bytes)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  const float* sHB = reinterpret_cast<const float*>(sW + img.off_f32) + 128 + 128 + 3 * 128 + 4;   // REF head biases [48]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time setup: barriers, weights -> smem, TMEM -------------------------------------------
  // The 86 KB operand image is copied by the TMA engine (cp.async.bulk, SASS UBLKCP): one thread issues six bulk copies
  // and goes on; nobody spends load / store instructions on it, and the copy overlaps the TMEM allocation and the
  // register re-split.  Everybody waits for the transaction bytes on w_bar before the first use.
  if (tid == 0) {
    mbar_init(&full[0], kRows / 8);      // one arrival per 8-row pass
    mbar_init(&full[1], kRows / 8);
    for (int i = 2; i < 8; ++i) mbar_init(&bars[i], 1);
    mbar_init(w_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(w_bar)), "r"(img.bytes_fwd) : "memory");
    constexpr uint32_t kChunk = 16384;
    const uint8_t* src = reinterpret_cast<const uint8_t*>(P.m.tc_weights);
    for (uint32_t off = 0; off < img.bytes_fwd; off += kChunk) {
      const uint32_t sz = min(kChunk, img.bytes_fwd - off);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(sW + off)), "l"(src + off), "r"(sz), "r"(smem_u32(w_bar)) : "memory");
    }
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_slot;
  mbar_wait(w_bar, 0);

  constexpr uint32_t LBO_A = kLboA0, LBO_B0 = NH * 16, LBO_B = 128 * 16, LBO_B3 = 16 * 16, SBO = 128;
  constexpr uint32_t IDESC_NH = instr_desc(128, NH, H16), IDESC_N128 = instr_desc(128, 128, H16), IDESC_N16 = instr_desc(128, 16, H16);
  constexpr uint32_t A0_STAGE = (K0 / 8) * kLboA0;

  const uint32_t n_ent = min(*P.ws.n_entries, P.ws.cap);      // bounded workspace: entries behind the capacity were not stored
  const uint32_t tile_rows = balanced_tile_rows(n_ent, kGroups * gridDim.x);   // 128, or fewer for launches of a few rounds
  const uint32_t n_tiles = (n_ent + tile_rows - 1) / tile_rows;

  if (warp >= kMlpWarps2) {
    // =============================== gather group ================================================
    constexpr int kLaunchRegs = (65536 / kThreads) / 8 * 8;
    if (RegSplit<NGW>::kGather <= kLaunchRegs) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(RegSplit<NGW>::kGather));
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(RegSplit<NGW>::kGather));
    const int gw = warp - kMlpWarps2;
    // A tile is 16 passes of 8 rows (4 lanes per entry); pass number p = 16 * it + pass of the CTA's tile sequence goes to
    // gather warp p % NGW, so any warp count divides the work evenly over a few tiles.  Every pass arrives on full[stage]
    // (count 16).  Entry coordinates: written by k_march with the coordinates it marched; read once: evict-first.
    // (Fetching them one pass ahead measured SLOWER, three times: held in registers 2.25 vs 2.07 ms per frame and 1.74 vs 1.47
    //  with the mixed-precision gather; by cp.async into shared memory, no register held, 1.64 vs 1.45 -- warps that run ahead
    //  of their neighbours lose the L1 hits they share with them; profiles/r02_notes.txt B, N, S.)
    constexpr uint32_t kPasses = kRows / 8;
    const uint32_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
#pragma unroll 1
    for (uint32_t p = gw; p < my_tiles * kPasses; p += NGW) {
      const uint32_t it = p / kPasses, pass = p % kPasses;
      const uint32_t s = it & 1u, use = it >> 1;
      mbar_wait(&empty[s], (use & 1u) ^ 1u);          // stage free (first use of a stage passes at once; the barrier cannot
                                                      // run a second phase ahead: that needs this very pass)
      const uint32_t row = pass * 8 + (lane >> 2);
      const uint32_t e = row < tile_rows ? (blockIdx.x + it * gridDim.x) * tile_rows + row : n_ent;      // dead rows of a short tile
      uint8_t* arow = sA0 + s * A0_STAGE + row * 16;
#ifdef TVM_EXP_NOGATHER
      gather_row<CA, PB16, H16>(P, arow, false, make_float4(0.0f, 0.0f, 0.0f, 0.0f), lane & 3);
#else
      const float4 uw = e < n_ent ? __ldcs(P.ws.ent_u + e) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#ifdef TVM_APP2_UNPACK
      gather_row<CA, PB16, H16>(P, arow, e < n_ent, uw, lane & 3);
#else
      if (PB16) gather_row_mixed<CA, H16>(P, arow, e < n_ent, uw, lane);
      else gather_row<CA, PB16, H16>(P, arow, e < n_ent, uw, lane & 3);
#endif
#endif
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
    }
  } else {
    // =============================== MLP groups ====================================================
    constexpr int kLaunchRegs = (65536 / kThreads) / 8 * 8;
    if (RegSplit<NGW>::kMlp >= kLaunchRegs) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(RegSplit<NGW>::kMlp));
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(RegSplit<NGW>::kMlp));
    const int g = warp >> 2;                                // group 0 / 1
    const int row = tid & 127;                              // tile row = TMEM lane
    const bool leader = row == 0;
    const uint32_t tg = tmem + (uint32_t)g * kGroupCols;
    const uint32_t lane_addr = tg + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t aA0 = smem_u32(sA0 + g * A0_STAGE);
    const uint32_t aB0 = smem_u32(sW + img.off_b0), aB1 = smem_u32(sW + img.off_b1), aB2 = smem_u32(sW + img.off_b2);
    const uint32_t aB2x = smem_u32(sW + img.off_b2x), aB3 = smem_u32(sW + img.off_b3);
    uint64_t* mma_bar = &mma_bars[g];
    uint64_t* bas_bar = &bas_bars[g];
    uint32_t mma_phase = 0, j = 0;                          // j: tiles this group has started
    float pen_acc = 0.0f;
    // EV (TVM_EVAL_ONLY): the fixed-point terms of the row's PREVIOUS tile, added to their ray while GEMM1 of the next tile runs
    // (issued in the last epilogue the three REDs of a row sit on the tile's critical path: 1.569 vs 1.553 ms per frame; summing
    // the rows of a warp first -- REDUX when they all belong to one ray, the rule in foggy scenes -- changes nothing: the cost
    // of the mode is not the number of atomics, profiles/r02_notes.txt AD)
    uint32_t pend_ray = 0xffffffffu, pend0 = 0, pend1 = 0, pend2 = 0;
    auto flush_pending = [&]() {
      if (pend_ray != 0xffffffffu) {
        uint32_t* a = fix_sums(P.ws) + 3 * (size_t)pend_ray;
        atomicAdd(a, pend0); atomicAdd(a + 1, pend1); atomicAdd(a + 2, pend2);         // results unused: RED.E.ADD
        pend_ray = 0xffffffffu;
      }
    };

    // GEMM0 of this group's tile number jj (leader only): feat = A0[g] . basis^T -> BAS columns
    auto issue_gemm0 = [&](uint32_t jj) {
      mbar_wait(&full[g], jj & 1u);
      fence_after();
#pragma unroll
      for (int k = 0; k < K0 / 16; ++k)
        umma_bf16(tg + kColBas, smem_desc(aA0 + k * 2 * LBO_A, LBO_A, SBO), smem_desc(aB0 + k * 2 * LBO_B0, LBO_B0, SBO), IDESC_NH, k > 0);
      umma_commit(&empty[g]);     // the stage may be refilled once these MMAs have read it
      umma_commit(bas_bar);
    };
    const uint32_t stride = 2u * gridDim.x;
    uint32_t tile = blockIdx.x + (uint32_t)g * gridDim.x;
    if (leader && tile < n_tiles) issue_gemm0(0);

    for (; tile < n_tiles; tile += stride, ++j) {
      const uint32_t e = (uint32_t)row < tile_rows ? tile * tile_rows + row : n_ent;                     // dead rows of a short tile
      float dir[3] = {0.0f, 0.0f, 0.0f};
      uint32_t ray = 0;
      float wgt = 0.0f;
      if (e < n_ent) {
        ray = __ldcs(&P.ws.ent[e].x);               // read once: evict-first, the L1 belongs to the gather's texels
        if (EV) wgt = __ldcs(P.ws.ent_w + e);
        dir[0] = P.rays[6 * (size_t)ray + 3];
        dir[1] = P.rays[6 * (size_t)ray + 4];
        dir[2] = P.rays[6 * (size_t)ray + 5];
      }
      mbar_wait(bas_bar, j & 1u);
      fence_after();
#ifdef TVM_EXP_NOMLP
      fence_before();
      group_sync(g);
      if (leader && tile + stride < n_tiles) { fence_after(); issue_gemm0(j + 1); }
      continue;
#endif
      float rgb_d0 = 0.0f, rgb_d1 = 0.0f, rgb_d2 = 0.0f, tint = 1.0f;
      {
        // ---- epi0: features -> [feat, view, sin/cos PE] as 16-bit pairs into the TMEM A operand (tensorBase.py:76-83, 9-15)
        float x[32];
        tmem_ld32(lane_addr + kColBas, x);
        float ndot = 0.0f;
        if (REF) {
          // REFTensoRF.py:216-232: heads -> unit normal, reflected direction, -dot, diffuse colour, tint
          float hx[16];
          tmem_ld16(lane_addr + kColBas + 32, hx);
          float nx = x[APP_DIM] + sHB[APP_DIM], ny = x[APP_DIM + 1] + sHB[APP_DIM + 1], nz = x[APP_DIM + 2] + sHB[APP_DIM + 2];
          auto head = [&](int o) { return (o < 32 ? x[o] : hx[o - 32]) + sHB[o]; };
          rgb_d0 = head(APP_DIM + 3); rgb_d1 = head(APP_DIM + 4); rgb_d2 = head(APP_DIM + 5);
          tint = fmaxf(head(APP_DIM + 6), 0.0f);
          const float inv = rsqrtf(fmaxf(nx * nx + ny * ny + nz * nz, 1e-30f));
          nx *= inv; ny *= inv; nz *= inv;
          const float dx = -dir[0], dy = -dir[1], dz = -dir[2];
          const float dot = dx * nx + dy * ny + dz * nz;
          dir[0] = 2.0f * dot * nx - dx;
          dir[1] = 2.0f * dot * ny - dy;
          dir[2] = 2.0f * dot * nz - dz;
          ndot = -dot;
          if (e < n_ent) {
            const float pen = fmaxf(-dot, 0.0f);
            P.ws.ent_pen[e] = pen * pen;
            if (P.aux.penalty) pen_acc = fmaf(P.ws.ent_w[e], pen * pen, pen_acc);     // one atomic per warp at the end of the CTA
          }
        }
        float s1[APP_DIM + 3], c1[APP_DIM + 3];
#pragma unroll
        for (int o = 0; o < APP_DIM; ++o) __sincosf(x[o], &s1[o], &c1[o]);
#pragma unroll
        for (int o = 0; o < 3; ++o) __sincosf(dir[o], &s1[APP_DIM + o], &c1[APP_DIM + o]);
        auto column = [&](int cc) -> float {
          constexpr int PF = APP_DIM + 3, NF = FEA_PE * APP_DIM, PV = PF + 2 * NF, NV = VIEW_PE * 3;
          if (REF && cc == 0) return ndot;
          const int c = cc - C0;
          if (c < APP_DIM) return x[c];
          if (c < PF) return dir[c - APP_DIM];
          if (c < PF + NF) { const int o = (c - PF) >> 1; return ((c - PF) & 1) ? 2.0f * s1[o] * c1[o] : s1[o]; }
          if (c < PV) { const int o = (c - PF - NF) >> 1; return ((c - PF - NF) & 1) ? 1.0f - 2.0f * s1[o] * s1[o] : c1[o]; }
          if (c < PV + NV) { const int o = APP_DIM + ((c - PV) >> 1); return ((c - PV) & 1) ? 2.0f * s1[o] * c1[o] : s1[o]; }
          if (c < PV + 2 * NV) { const int o = APP_DIM + ((c - PV - NV) >> 1); return ((c - PV - NV) & 1) ? 1.0f - 2.0f * s1[o] * s1[o] : c1[o]; }
          if (cc == IN_C) return 1.0f;                 // constant-one column: row IN_C of the W1 operand is b1
          return 0.0f;
        };
#pragma unroll
        for (int blk = 0; blk < K1 / 32; ++blk) {
          uint32_t w[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) w[i] = pack16<H16>(column(blk * 32 + 2 * i), column(blk * 32 + 2 * i + 1));
          tmem_st16(lane_addr + kColA + blk * 16, w);
        }
      }
      tmem_st_wait();
      fence_before();
      group_sync(g);
      // ---- GEMM1: A1 . W1^T (b1 rides on the constant-one column); then GEMM0 of this group's next tile ---------------
      if (leader) {
        fence_after();
#pragma unroll
        for (int k = 0; k < K1 / 16; ++k)
          umma_ts(tg + kColD, tg + kColA + k * 8, smem_desc(aB1 + k * 2 * LBO_B, LBO_B, SBO), IDESC_N128, k > 0);
        umma_commit(mma_bar);
        if (tile + stride < n_tiles) issue_gemm0(j + 1);     // BAS was drained by epi0 of every row before the group barrier
      }
      if (EV) flush_pending();      // under GEMM1
      mbar_wait(mma_bar, mma_phase);
      mma_phase ^= 1;
      fence_after();
      // ---- epi1: ReLU -> A2 (K 0..127), constant one at K = 128 (row 128 of the W2 / W3 operands is b2 / b3), zeros to 143 ----
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        float y[32];
        tmem_ld32(lane_addr + kColD + cb * 32, y);
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = pack16_relu<H16>(y[2 * i], y[2 * i + 1]);
        tmem_st16(lane_addr + kColA + cb * 16, w);
      }
      {
        const uint32_t w[8] = {H16 ? 0x00003c00u : 0x00003f80u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        tmem_st8(lane_addr + kColA + 64, w);
      }
      tmem_st_wait();
      fence_before();
      group_sync(g);
      // ---- GEMM2: [A2 | 1] . [W2^T; b2] -------------------------------------------------------------------------------------
      if (leader) {
        fence_after();
#pragma unroll
        for (int k = 0; k < 128 / 16; ++k)
          umma_ts(tg + kColD, tg + kColA + k * 8, smem_desc(aB2 + k * 2 * LBO_B, LBO_B, SBO), IDESC_N128, k > 0);
        umma_ts(tg + kColD, tg + kColA + 64, smem_desc(aB2x, LBO_B, SBO), IDESC_N128, 1u);
        umma_commit(mma_bar);
      }
      mbar_wait(mma_bar, mma_phase);
      mma_phase ^= 1;
      fence_after();
      // ---- epi2: ReLU -> A3 (the constant-one column of A2 stays) -----------------------------------------------------------
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        float y[32];
        tmem_ld32(lane_addr + kColD + cb * 32, y);
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = pack16_relu<H16>(y[2 * i], y[2 * i + 1]);
        tmem_st16(lane_addr + kColA + cb * 16, w);
      }
      tmem_st_wait();
      fence_before();
      group_sync(g);
      // ---- GEMM3: [A3 | 1] . [W3^T; b3] (N = 16, 3 real columns) -> D columns 0..15 ------------------------------------------
      if (leader) {
        fence_after();
#pragma unroll
        for (int k = 0; k < 144 / 16; ++k)
          umma_ts(tg + kColD, tg + kColA + k * 8, smem_desc(aB3 + k * 2 * LBO_B3, LBO_B3, SBO), IDESC_N16, k > 0);
        umma_commit(mma_bar);
      }
      mbar_wait(mma_bar, mma_phase);
      mma_phase ^= 1;
      fence_after();
      {
        float o[16];
        tmem_ld16(lane_addr + kColD, o);
        // REF: rgb = tint * clamp(rgb_s, 0) + rgb_d (REFTensoRF.py:232); VM: tint = 1, rgb_d = 0
        const float c0 = tint / (1.0f + __expf(-o[0])) + rgb_d0, c1 = tint / (1.0f + __expf(-o[1])) + rgb_d1;
        const float c2 = tint / (1.0f + __expf(-o[2])) + rgb_d2;
        if (EV) {
          // composite as we go (tvmrender.h: TVM_EVAL_ONLY): w * rgb of this row in fixed point, added to its ray's sums one
          // tile later (flush_pending).  (Summing each run of rows that belong to one ray with a segmented shuffle reduction
          // first -- one RED per run -- measured slower than a RED per row, 1.584 vs 1.567 ms per frame.)
          if (e < n_ent) {
            pend_ray = ray;
            pend0 = __float2uint_rn(wgt * c0 * kFixScale);
            pend1 = __float2uint_rn(wgt * c1 * kFixScale);
            pend2 = __float2uint_rn(wgt * c2 * kFixScale);
          }
        } else if (e < n_ent) {
          P.ws.ent_rgb[(size_t)e * 3 + 0] = c0;
          P.ws.ent_rgb[(size_t)e * 3 + 1] = c1;
          P.ws.ent_rgb[(size_t)e * 3 + 2] = c2;
        }
      }
      // no barrier here: the next tile's epi0 touches this thread's own TMEM lane only (A after GEMM3 has completed, BAS), and
      // GEMM1 of the next tile -- the next writer of D -- is issued behind the group barrier that follows epi0
      fence_before();
    }
    if (EV) flush_pending();        // the last tile's rows
    if (REF && P.aux.penalty) {
      pen_acc = warp_sum(pen_acc);
      if (lane == 0 && pen_acc != 0.0f) atomicAdd(P.aux.penalty, pen_acc);
    }
  }

  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int NGW, bool EV>
static int launch(const FwdParams& P, int num_sms, cudaStream_t stream, bool ref, bool pb16, bool h16, size_t smem) {
  void (*kern)(const FwdParams);
  if (EV) {           // VM only (forward_impl clears the flag otherwise)
    if (h16) kern = pb16 ? k_app_tc2<48, 27, 2, 2, false, true, true, NGW, EV> : k_app_tc2<48, 27, 2, 2, false, false, true, NGW, EV>;
    else kern = pb16 ? k_app_tc2<48, 27, 2, 2, false, true, false, NGW, EV> : k_app_tc2<48, 27, 2, 2, false, false, false, NGW, EV>;
  } else if (h16)
    kern = ref ? (pb16 ? k_app_tc2<48, 27, 2, 2, true, true, true, NGW, false> : k_app_tc2<48, 27, 2, 2, true, false, true, NGW, false>)
               : (pb16 ? k_app_tc2<48, 27, 2, 2, false, true, true, NGW, false> : k_app_tc2<48, 27, 2, 2, false, false, true, NGW, false>);
  else
    kern = ref ? (pb16 ? k_app_tc2<48, 27, 2, 2, true, true, false, NGW, false> : k_app_tc2<48, 27, 2, 2, true, false, false, NGW, false>)
               : (pb16 ? k_app_tc2<48, 27, 2, 2, false, true, false, NGW, false> : k_app_tc2<48, 27, 2, 2, false, false, false, NGW, false>);
  TVM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<num_sms, (kMlpWarps2 + NGW) * 32, smem, stream>>>(P);
  TVM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace app2

#ifndef TVM_APP2_GATHER_WARPS
#define TVM_APP2_GATHER_WARPS 16
#endif

// caller (launch_app_tc, tvm_mlp_tc.cu) has validated mode, shape and tc_weights
int launch_app_tc2(const FwdParams& P, int num_sms, cudaStream_t stream) {
  const bool h16 = (P.flags & TVM_MLP_MASK) == TVM_MLP_FP16;
  const bool ref = P.m.variant == TVM_VARIANT_REF;
  const Image img(P.m.n_app, P.in_mlp_c, head_ld(P.m));
  const size_t smem = ((img.bytes_fwd + 1023) & ~1023u) + 2 * (size_t)(img.K0 / 8) * app2::kLboA0 + 128 + 1024;
  const bool pb16 = P.m.app_plane_pair[0] && P.m.app_plane_pair[1] && P.m.app_plane_pair[2] && P.m.app_line_pair[0] &&
                    P.m.app_line_pair[1] && P.m.app_line_pair[2];
  if ((P.flags & TVM_EVAL_ONLY) && !ref) return app2::launch<TVM_APP2_GATHER_WARPS, true>(P, num_sms, stream, ref, pb16, h16, smem);
  return app2::launch<TVM_APP2_GATHER_WARPS, false>(P, num_sms, stream, ref, pb16, h16, smem);
}

}  // namespace tvm
