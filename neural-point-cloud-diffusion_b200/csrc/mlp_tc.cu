// G1/G2/M1/A1/M2/M3 on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only -- kernel generation 2.
// Same contract as mlp_simt.cu (reference: fields/aggregators/mlp.py:69-88,119-121, fields/mlp.py:38-72, fields/field.py:126-141).
//
// Precision: the reference runs TRUE fp32 GEMMs (no TF32, train_pointnerf.py:16-17) and the parity bar is 1e-4 on RGB after ten
// chained layers, which single-pass TF32/BF16 cannot meet.  Every operand is therefore split into two fp16 halves
// (x = hi + lo, 22 significant bits) and each layer is accumulated in fp32 TMEM from THREE tcgen05.mma.kind::f16 products
// (hi*hi + lo*hi + hi*lo; the dropped lo*lo term is 2^-22 relative).  Weights are pre-scaled by a power of two per layer (exact)
// so the lo halves stay out of the fp16 subnormal range; the inverse scale is folded into the epilogue FMA.
//
// Kernel shape (one persistent CTA per SM, 352 threads, warp-specialised):
//   warp 0      : weight producer.  Streams pre-swizzled 32 KB weight tiles (256 out x 64 k, hi or lo) from L2 into a 3-stage
//                 ring with cp.async.bulk (UBLKCP) completing on mbarriers.
//   warp 1      : allocates all 512 TMEM columns (two 128x256 fp32 accumulators), then one elected lane issues tcgen05.mma
//                 (M=128, N=256, K=16, cta_group::1), A and B from shared memory (K-major SWIZZLE_128B descriptors).
//   warps 2..9  : 256 prologue/epilogue threads, two per tile row.  The epilogue of layer l runs in 32-column chunks: tcgen05.ld,
//                 scale+bias+LeakyReLU, fp16 hi/lo split, store IN PLACE over the layer's own (already consumed) A operand, and
//                 publish each 64-column K-block on its own mbarrier -- the MMAs of layer l+1 (into the OTHER accumulator) start
//                 as soon as K-block 0 is there, so the tensor pipe only idles for one chunk per layer.
//   warp 10     : heads mode only: bulk-copies the next tile's pre-split A operand (written by the pair kernel) into the
//                 K-blocks the last layer's MMAs have released.
//   warps 11..14: inference pair mode only (ncu on the version where the epilogue threads did this between their layer epilogues:
//                 the tensor pipe was busy 48 % of the time, the 8 epilogue warps were the serial resource): one thread per row
//                 of the NEXT tile gathers the pair's point, computes the positional encoding and stores the layer-0 operand
//                 as soon as layer 3's MMAs release K-blocks 0 / 1; the same warps then take a third of the segmented-sum
//                 tasks of the aggregation epilogue.
// Inference with the "f16 + e4m3 x 2" operand scheme (tc_ptx.cuh) runs the kTS instantiations: every layer whose A operand is the
// previous layer's epilogue output reads it from TENSOR MEMORY (the epilogue converts its accumulator in place, see the comment at
// k_field_tc); shared memory then holds the weight ring, the gathered layer-0 input / the heads' tile image, and the aggregation
// staging.  The training modes and the "f16x3" / one-correction schemes keep every operand in shared memory as described above.
// Pair mode packs (sample, neighbour) pairs DENSELY: a tile is a maximal run of whole samples whose pairs fit its 128 rows (greedy,
// k_tile_walk below: ~125 used rows per tile), instead of 8 slots per sample (21 % padding at 6.3 neighbours per sample).
// The next tile's gather + positional encoding is computed by the epilogue threads while the tensor pipe works on layers 1..2
// and is stored during layer 3 into the K-blocks that layer has already consumed.
#include <cub/cub.cuh>
#include "npcd_b200.h"
#include "tc_ptx.cuh"

namespace npcd {
namespace tc {

constexpr int kThreadsTc = 352;
constexpr int kThreadsTcTrainHeads = 384;  // + warp 11: stash copies (warp 10 is the first-operand loader in heads mode)
constexpr int kThreadsTcPro = 480;         // inference pair mode: + warps 11..14, the input (gather / posenc) warps
constexpr int kProThreads = 128;
template <int kMode>
constexpr int threads_for() {
  return kMode == 0 /* MODE_PAIR */ ? kThreadsTcPro : (kMode == 4 /* MODE_HEADS_TRAIN */ ? kThreadsTcTrainHeads : kThreadsTc);
}
constexpr int kEpiThreads = 256;
constexpr int kTileBytesW = 256 * 128;        // one K-block of 256 output rows
constexpr int kStages = 3;
constexpr int kSmemA = 4 * 2 * kTileBytesA;   // 4 K-blocks x (hi, lo) = 128 KB; K-blocks 2..3 double as fp32 aggregation staging
constexpr int kSmemW = kStages * kTileBytesW; // 96 KB
constexpr int kSmemMisc = 3072;
constexpr int kSmemTotal = kSmemA + kSmemW + kSmemMisc;  // 232448 = the 227 KB per-CTA maximum
constexpr uint32_t kIdesc = make_idesc(128, 256);  // D=f32, A=B=f16, K-major, N=256, M=128
constexpr int kTmemCols = 512;
constexpr int kPackRows = 121;                // minimum rows of a greedy dense tile that is not the last of its block
constexpr int kImgTileBytes = 4 * 2 * kTileBytesA;  // one 128-sample tile of the pre-split [S,256] operand image (128 KB)

enum Epi { EPI_ACT = 0, EPI_LINEAR = 1, EPI_AGG = 2, EPI_DOT1 = 3, EPI_DOT3 = 4, EPI_DUMP = 5 };
enum Mode { MODE_PAIR = 0, MODE_HEADS = 1, MODE_PROBE = 2, MODE_PAIR_TRAIN = 3, MODE_HEADS_TRAIN = 4 };

// misc shared-memory layout (byte offsets)
constexpr int kOffBars = 0;       // 24 mbarriers
constexpr int kOffTmem = 192;
constexpr int kOffInfo = 208;     // int[2][4]: s_begin, n_samp, n_rows
constexpr int kOffEpiBusy = 240;  // int: the epilogue warps are inside a layer epilogue (the input warps yield issue slots, kYield)
constexpr int kOffWts = 256;      // float[2][128] raw inverse-distance weights           (pair)
constexpr int kOffRowSamp = 1280; // u8[2][128] row -> sample-in-tile                      (pair)
constexpr int kOffSampRow = 1536; // u8[2][128] sample-in-tile -> first row                (pair)
constexpr int kOffSampCnt = 1792; // u8[2][128] sample-in-tile -> rows                     (pair)
constexpr int kOffPart = 256;     // float[128][4] cross-half partial dot products, 2048 B  (heads; aliases the pair arrays)
// (the heads' partial-product array ends at 2304: every barrier placed after the first block must sit beyond it -- with the stash
// barriers at 2048 the dot epilogues of rows 112..113 used to write their partial sums over them)
constexpr int kOffStashBars = 2304;  // 4 mbarriers: K-block kb of the A operand has been copied to the training stash (+ kBarARdy2)

// barrier indices
constexpr int kBarWFull = 0, kBarWEmpty = 3, kBarARdy = 6, kBarA0Rdy = 10, kBarAFree = 14, kBarAccRdy = 18, kBarAccFree = 20;
// 2-SM mode (cta_group::2, inference pair / heads kernels): the leader CTA of a pair issues every MMA for both CTAs' tiles
// (M = 256) and each CTA holds HALF of every weight tile, so the 96 KB weight area is a ring of SIX 16 KB half-tiles.
// Why: the timeline of the 1-SM version shows a layer's 32 MMAs taking ~5600 cycles instead of 4096 -- shared-memory bandwidth:
// 96 B/clk of MMA operand reads (4 KB of A + 8 KB of B per 128 cycles) + 64 B/clk of weight refill + the epilogue's stores exceed
// the 128 B/clk an SM has.  With half of B per CTA the MMAs read 64 B/clk and the refill writes 32 B/clk.
// MEASURED (B200, round 2, profiles/r2_two_sm_experiment.md): bit-identical results, but NOT faster yet -- 153.5 ms per 251-view
// step against 141.9 ms for the 1-SM kernel: the cross-CTA hand-offs (remote arrives of the peer's epilogue / input warps, the
// relay of the peer's weight-full barrier) lengthen the epilogue -> MMA chain by more than the MMA phase gains (with cluster-scope
// release / acquire on those hand-offs it was 174 ms).  Build switch, OFF by default.
#ifndef NPCD_TC_CLUSTER
#define NPCD_TC_CLUSTER 1
#endif
#ifndef NPCD_TC_2SM
#define NPCD_TC_2SM 0
#endif
template <int kMode>
constexpr bool two_sm() {
  return NPCD_TC_2SM != 0 && NPCD_TC_CLUSTER != 0 && (kMode == 0 /* MODE_PAIR */ || kMode == 1 /* MODE_HEADS */);
}
constexpr int kStages2 = 6;
constexpr int kOffBars2 = 2368;  // 2-SM barrier block in the misc area (22 barriers, ends at 2544 of the 3072 bytes)
static_assert(kOffPart + 128 * 4 * 4 <= kOffStashBars && kOffStashBars + 5 * 8 <= kOffBars2 && kOffBars2 + 22 * 8 <= kSmemMisc,
              "misc shared-memory layout");
constexpr int kBar2WFull = kOffBars2 / 8, kBar2WEmpty = kBar2WFull + 6, kBar2WPeer = kBar2WFull + 12, kBar2A0Peer = kBar2WFull + 18;
constexpr int kBarStash = kOffStashBars / 8;
constexpr int kBarARdy2 = kBarStash + 4;  // second half (columns 16-31 / 48-63) of K-block 0 published (inference: split first block)

struct Layer {
  const uint8_t* w;  // packed tiles: for each K-block: hi tile (32 KB) then lo tile (32 KB)
  float inv_scale;   // weights were multiplied by 1/inv_scale (a power of two) when packed
  int ksteps;        // K / 16
  int epi;
};

// Passed by value (constant bank): biases and the two narrow output layers are read with warp-uniform constant loads.
struct Params {
  Layer layers[6];
  float bias[6][256];
  float shape_out_w[256];
  float chan_out_w[3][256];
  float shape_out_b;
  float chan_out_b[3];
  int n_layers;
  int layer_ofs;  // heads: 1 when local_field.8 is folded into shape_net.0 / channel_net.0 (layers[0] = shape_net.0'), else 0
  // pair
  const int* nbr_idx;
  const float4* sample_pos;
  const float* kp_pos;
  const float* kp_feat;
  const int* pair_off;    // [S+1] exclusive scan of the neighbour counts
  const int* tile_start;  // [n_tiles+1] first sample of every dense tile
  const int* n_tiles_dev;
  uint8_t* img;           // pre-split [S,256] operand image: pair mode writes it, heads / probe mode read it
  // heads
  float4* rgbs;
  float* feat_out;  // optional [S,256] fp32 (heads: local_field output; probe: layer output)
  const long long* n_samples_dev;
  long long capacity;
  int* error_flag;
  // training stash (MODE_PAIR_TRAIN): everything the fused backward (pair_bwd_tc.cu) and the weight-gradient GEMMs need
  uint8_t* stash_x[4];      // operand images of the layer inputs X_0 (96 columns, 2 K-blocks) and X_1..X_3 (4 K-blocks), per tile
  uint32_t* stash_mask[4];  // [tile][128 rows][8] sign bits (y > 0) of the outputs of layers 0..3 (LeakyReLU derivative)
  float* stash_wn;          // [tile][128] normalised inverse-distance weight of every row
  int* stash_idx;           // [tile][128] global point index of every row (-1 = padding)
  int* stash_samp;          // [tile][128] sample index of every row (-1 = padding)
  // training stash of the heads stage (MODE_HEADS_TRAIN), per 128-sample tile
  uint8_t* hstash_x[6];      // operand images of F (local_field.8 output), C1, C2, C3 (channel_net hidden 1..3), C4, H (shape hidden)
  uint32_t* hstash_mask[5];  // sign bits of H, C1, C2, C3, C4
  long long* timeline;       // development aid (npcd_debug_set_timeline): clock64() of CTA 0's phase boundaries, [tile][32] events
  int no_wcorr;              // f16 + e4m3 scheme with ONE correction product: the weight-rounding correction hi8 x Wlo8 is not issued
};

constexpr int kTimelineTiles = 64;
// event e of the `it`-th tile of CTA 0 (one designated thread per role calls this; a null pointer costs one uniform branch)
#define NPCD_TL(it, e)                                                                                        \
  do {                                                                                                        \
    if (P.timeline && blockIdx.x == 0 && (it) < kTimelineTiles) P.timeline[(it) * 32 + (e)] = clock64();      \
  } while (0)
#define NPCD_TLV(it, e, val)                                                                                  \
  do {                                                                                                        \
    if (P.timeline && blockIdx.x == 0 && (it) < kTimelineTiles) P.timeline[(it) * 32 + (e)] = (val);          \
  } while (0)

// One chunk (32 accumulator columns starting at c0) of an ACT / LINEAR epilogue: y = [lrelu](acc * inv + b) -> fp16 hi/lo ->
// K-block (c0 >> 6) of the A operand, 16-byte chunks (c0 & 63) / 8 .. +3.
// slope = 0.01 (LeakyReLU) or 1 (linear layer: max(y, y) = y).
// kG0 .. kG1: the 8-column groups of the chunk to process (0 .. 4 = all; the first K-block of a layer is published in two halves)
template <bool kF8 = false, int kG0 = 0, int kG1 = 4>
__device__ __forceinline__ void epi_chunk_store(const Params& P, int l, float inv, float slope, const uint32_t (&v)[32], int c0,
                                                uint8_t* sA, uint32_t rowbase, int x7, float* feat_row, uint32_t* mask_out = nullptr) {
  uint32_t mbits = 0u;
  uint8_t* kb_base = sA + (c0 >> 6) * (2 * kTileBytesA) + rowbase;
  if (NPCD_EXP_NOEPI) {  // timing-only ablation: no arithmetic, no stores -- the chain is TMEM load -> publish
    if (v[kG0 * 8] == 0x12345678u) *reinterpret_cast<uint32_t*>(kb_base) = v[kG0 * 8 + 1];
    return;
  }
  const int c16_0 = (c0 & 63) >> 3;
  const uint64_t inv2 = pack2(inv, inv), slope2 = pack2(slope, slope);
  static_assert(kG0 % 2 == 0 && kG1 % 2 == 0, "8-bit tile: groups are stored in pairs");
  uint2 lo8_even = make_uint2(0u, 0u), hi8_even = make_uint2(0u, 0u);
#pragma unroll
  for (int g = kG0; g < kG1; ++g) {
    const float4 b0 = NPCD_EXP_NOBIAS ? make_float4(0.1f, 0.2f, 0.3f, 0.4f) : *reinterpret_cast<const float4*>(&P.bias[l][c0 + g * 8]);
    const float4 b1 = NPCD_EXP_NOBIAS ? make_float4(0.1f, 0.2f, 0.3f, 0.4f) : *reinterpret_cast<const float4*>(&P.bias[l][c0 + g * 8 + 4]);
    float y[8];
    act2(v[g * 8 + 0], v[g * 8 + 1], inv2, b0.x, b0.y, slope2, y[0], y[1]);
    act2(v[g * 8 + 2], v[g * 8 + 3], inv2, b0.z, b0.w, slope2, y[2], y[3]);
    act2(v[g * 8 + 4], v[g * 8 + 5], inv2, b1.x, b1.y, slope2, y[4], y[5]);
    act2(v[g * 8 + 6], v[g * 8 + 7], inv2, b1.z, b1.w, slope2, y[6], y[7]);
    if (feat_row) {
      const float fm = kF8 ? 1.0f / kF8ActScale : 1.0f;  // the f8 scheme carries activations times kF8ActScale
      *reinterpret_cast<float4*>(feat_row + c0 + g * 8) = make_float4(y[0] * fm, y[1] * fm, y[2] * fm, y[3] * fm);
      *reinterpret_cast<float4*>(feat_row + c0 + g * 8 + 4) = make_float4(y[4] * fm, y[5] * fm, y[6] * fm, y[7] * fm);
    }
    if (mask_out) {
#pragma unroll
      for (int j = 0; j < 8; ++j) mbits |= (y[j] > 0.f ? 1u : 0u) << (g * 8 + j);
    }
    uint8_t* p = kb_base + (((c16_0 + g) ^ x7) << 4);
    if (!kF8) {
      uint4 hi, lo;
      split8(y, hi, lo);
      *reinterpret_cast<uint4*>(p) = hi;
      *reinterpret_cast<uint4*>(p + kTileBytesA) = lo;
    } else {
      uint4 hi;
      uint2 lo8, hi8;
      split8_f8(y, hi, lo8, hi8);
      *reinterpret_cast<uint4*>(p) = hi;
      const int c8 = c16_0 + g;  // 8-column group within the K-block
      // The lo8 / hi8 bytes of two neighbouring 8-column groups share a 16-byte chunk: one 16-byte store per 16 columns.  (ncu on
      // the version with 8-byte stores: 4 wavefronts per STS.64 instead of 2 -- the 32 rows of a warp put their 8 bytes into the
      // same half of every bank group -- 3.0e9 store bank conflicts per launch, 41 % of all shared-store wavefronts.)
      uint8_t* q = kb_base + kTileBytesA + ((f8_chunk(c8) ^ x7) << 4);
      uint8_t* r = kb_base + kTileBytesA + (((f8_chunk(c8) + 4) ^ x7) << 4);
      if (!(g & 1)) {
        lo8_even = lo8;
        hi8_even = hi8;
      } else if (!NPCD_EXP_NOSTS8) {
        *reinterpret_cast<uint4*>(q) = make_uint4(lo8_even.x, lo8_even.y, lo8.x, lo8.y);
        *reinterpret_cast<uint4*>(r) = make_uint4(hi8_even.x, hi8_even.y, hi8.x, hi8.y);
      } else if (lo8.x == 0x12345678u && hi8.y == 0x9abcdef0u) {  // keeps the conversions alive
        *reinterpret_cast<uint4*>(q) = make_uint4(lo8_even.x, lo8_even.y, lo8.x, lo8.y);
      }
    }
  }
  if (mask_out) *mask_out = mbits;
}

// The same epilogue arithmetic for the tensor-memory operand form (kTS): every 16 values of the chunk (groups 2 p, 2 p + 1) go back
// into the 16 accumulator columns they came from, t_chunk + 16 p: [fp16 x 16 (8 columns) | lo8 x 16 (4) | hi8 x 16 (4)].
template <int kG0 = 0, int kG1 = 4>
__device__ __forceinline__ void epi_chunk_ts(const Params& P, int l, float inv, float slope, const uint32_t (&v)[32], int c0,
                                             uint32_t t_chunk) {
  static_assert(kG0 % 2 == 0 && kG1 % 2 == 0, "pieces of 16 values");
  const uint64_t inv2 = pack2(inv, inv), slope2 = pack2(slope, slope);
#pragma unroll
  for (int p = kG0 / 2; p < kG1 / 2; ++p) {
    uint32_t w[16];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int g = 2 * p + h;
      const float4 b0 = *reinterpret_cast<const float4*>(&P.bias[l][c0 + g * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&P.bias[l][c0 + g * 8 + 4]);
      float y[8];
      act2(v[g * 8 + 0], v[g * 8 + 1], inv2, b0.x, b0.y, slope2, y[0], y[1]);
      act2(v[g * 8 + 2], v[g * 8 + 3], inv2, b0.z, b0.w, slope2, y[2], y[3]);
      act2(v[g * 8 + 4], v[g * 8 + 5], inv2, b1.x, b1.y, slope2, y[4], y[5]);
      act2(v[g * 8 + 6], v[g * 8 + 7], inv2, b1.z, b1.w, slope2, y[6], y[7]);
      uint4 hi;
      uint2 lo8, hi8;
      split8_f8(y, hi, lo8, hi8);
      w[4 * h + 0] = hi.x; w[4 * h + 1] = hi.y; w[4 * h + 2] = hi.z; w[4 * h + 3] = hi.w;
      w[8 + 2 * h] = lo8.x; w[8 + 2 * h + 1] = lo8.y;
      w[12 + 2 * h] = hi8.x; w[12 + 2 * h + 1] = hi8.y;
    }
    tmem_st16(t_chunk + 16u * (uint32_t)p, w);
  }
}

// Weight sharing across a 2-CTA cluster.  ncu on the single-CTA version: 878 GB per step of L2 -> SM reads, all of it the 0.90 MB of
// packed weights re-streamed for every 128-row tile -- energy that buys no FLOPs on a board sitting on its power cap.  Both CTAs of a
// cluster walk the same weight schedule, so CTA 0's producer issues ONE multicast bulk copy per ring stage that lands in both shared
// memories (and completes both full-barriers); a stage is reused once BOTH CTAs' MMAs have drained it (commit multicast to both
// empty-barriers, count 2).  A CTA that has one tile fewer than its partner keeps consuming the ring without issuing MMAs.
#ifndef NPCD_TC_CLUSTER
#define NPCD_TC_CLUSTER 1
#endif
constexpr bool kCluster = NPCD_TC_CLUSTER != 0;
#if NPCD_TC_CLUSTER
#define NPCD_TC_CLUSTER_DIMS __cluster_dims__(2, 1, 1)
#else
#define NPCD_TC_CLUSTER_DIMS
#endif

// ------------------------------------------------------------------------------------------------------------- kernel ----
// kTS (inference pair kernel, f16 + e4m3 x 2 scheme): the A operand of layers 1..3 lives in TENSOR MEMORY.  A layer epilogue
// converts its accumulator IN PLACE -- the 16 fp32 columns a thread has just loaded take the fp16 image (8 columns), the lo8 bytes
// (4 columns) and the hi8 bytes (4 columns) of those 16 values -- and the next layer's MMAs read that buffer as their A operand
// while accumulating into the other one (the trick of attention kernels that keep P over S).  Per 16 input features: one K16
// kind::f16 MMA on columns [0, 8) and ONE K32 kind::f8f6f4 MMA on columns [8, 16) = [lo8 | hi8] against weight rows [Whi8 | Wlo8]
// (weight format 2) -- both correction products in one instruction.  Shared memory then only feeds B (8 KB instead of 12 KB per
// M128 N256 step), the epilogues store nothing into shared memory, and K-blocks 0..1 are free for the next tile's layer-0 input as
// soon as layer 0's MMAs are done.  Layer 0 itself still reads its gathered input from shared memory.
template <int kMode, bool kF8 = false, bool kTS = false>
__global__ void NPCD_TC_CLUSTER_DIMS __launch_bounds__(threads_for<kMode>(), 1)
    k_field_tc(const __grid_constant__ Params P) {
  static_assert(!kTS || ((kMode == MODE_PAIR || kMode == MODE_HEADS) && kF8), "TS form: inference kernels, f8 scheme");
  // first layer whose A operand is the previous layer's epilogue output (and therefore, kTS, in tensor memory): layer 1 of the
  // pair MLP; channel_net.2 of the FOLDED heads (layers: shape_net.0', channel_net.0' -- both read the tile image in shared
  // memory --, channel_net.2, .4, .6)
  constexpr int kTsFirst = kMode == MODE_PAIR ? 1 : 2;
  constexpr bool kPro = kMode == MODE_PAIR;  // dedicated input warps (11..14)
  // inference: the epilogues publish K-block 0 of their output in two halves, so the next layer's MMAs start after 16 values per
  // thread instead of 32 (timeline of CTA 0: the first-block latency was ~1450 of the ~7100 cycles a layer takes)
  constexpr bool kSplit0 = kMode == MODE_PAIR || kMode == MODE_HEADS;
  static_assert(!kF8 || kMode == MODE_PAIR || kMode == MODE_HEADS || kMode == MODE_PROBE, "the f8 operand scheme is inference-only");
  constexpr bool kPair = kMode == MODE_PAIR || kMode == MODE_PAIR_TRAIN;
  constexpr bool kTrainP = kMode == MODE_PAIR_TRAIN;
  constexpr bool kTrainH = kMode == MODE_HEADS_TRAIN;
  constexpr bool kTrain = kTrainP || kTrainH;
  constexpr bool kHeads = kMode == MODE_HEADS || kMode == MODE_HEADS_TRAIN;
  constexpr bool k2 = two_sm<kMode>();
  constexpr bool kDynAgg = kPro && !k2 && !NPCD_EXP_NOREORDER;  // next tile's layer-0 epilogue interleaved with the aggregation
  // The input warps (one per scheduler, next to two epilogue warps) pause their gather / positional-encoding arithmetic while a
  // layer epilogue runs: timeline of the tensor-memory build -- the epilogue that overlapped it took 4 280 cycles instead of 3 050
  // and the MMAs of the layer behind it waited for operands (5 560 instead of 4 890 cycles per layer).
  constexpr bool kYield = kDynAgg && !NPCD_EXP_NOYIELD;
  constexpr int kSt = k2 ? kStages2 : kStages;                       // weight-ring stages
  constexpr int kStageBytes = k2 ? kTileBytesW / 2 : kTileBytesW;    // 2-SM: this CTA's half (128 of the 256 output rows) of a tile
  constexpr int kWF = k2 ? kBar2WFull : kBarWFull, kWE = k2 ? kBar2WEmpty : kBarWEmpty;
  constexpr uint32_t kIdescMma = k2 ? make_idesc(256, 256) : kIdesc;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sW = smem + kSmemA;
  uint8_t* misc = smem + kSmemA + kSmemW;
  const uint32_t bars = smem_u32(misc + kOffBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + kOffTmem);
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if ((smem_u32(smem) & 1023u) != 0u) {  // SWIZZLE_128B atoms need 1024-byte aligned tiles
    if (threadIdx.x == 0 && P.error_flag) atomicExch(P.error_flag, 1);
    return;
  }
  if (threadIdx.x == 0) {
    // 2-SM: the leader's issuer waits for the warps of BOTH CTAs (counts double, the peer arrives remotely); weight-full barriers are
    // local (each CTA's producer loads its half) + a relay barrier in the leader that the peer arrives on when its half has landed
    for (int i = 0; i < kSt; ++i) { mbar_init(bar(kWF + i), 1); mbar_init(bar(kWE + i), (kCluster && !k2) ? 2 : 1); }
    if (k2) {
      for (int i = 0; i < kSt; ++i) mbar_init(bar(kBar2WPeer + i), 1);
      for (int i = 0; i < 4; ++i) mbar_init(bar(kBar2A0Peer + i), 1);
    }
    // heads training: a K-block is free for the next tile's first operand once the last layer's MMAs AND the stash copy are done
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar(kBarARdy + i), k2 ? 16 : 8);
      mbar_init(bar(kBarA0Rdy + i), kPro ? (k2 ? 8 : 4) : 1);
      mbar_init(bar(kBarAFree + i), kTrainH ? 2 : 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(bar(kBarAccRdy + i), 1); mbar_init(bar(kBarAccFree + i), k2 ? 16 : 8); }
    for (int i = 0; i < 4; ++i) mbar_init(bar(kBarStash + i), 1);
    mbar_init(bar(kBarARdy2), k2 ? 16 : 8);
    *reinterpret_cast<volatile int*>(misc + kOffEpiBusy) = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (k2) tmem_alloc_2sm(smem_u32(tmem_slot), kTmemCols); else tmem_alloc(smem_u32(tmem_slot), kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  if (kCluster) cluster_sync_all();  // the partner's barriers are initialised before anything is multicast into them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long long S = min(*P.n_samples_dev, P.capacity);
  const int n_tiles = kPair ? *P.n_tiles_dev : (int)((S + 127) / 128);
  const int n_layers = P.n_layers;
  // weight-ring passes of this cluster = tiles of its even CTA (the odd one has as many or one fewer)
  const uint32_t cta_rank = kCluster ? cluster_ctarank() : 0u;
  const int first_cta = kCluster ? (int)(blockIdx.x & ~1u) : (int)blockIdx.x;
  const int n_pass = n_tiles > first_cta ? (n_tiles - first_cta + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  // 2-SM: both CTAs walk n_pass passes in lockstep; a CTA whose tile number runs past n_tiles processes a PHANTOM tile (no
  // samples, nothing stored) so that every barrier of the pair still gets its arrivals.
  // arrive on a barrier the MMA issuer waits on: 2-SM -> always the leader's copy (remote arrive from the peer)
  auto arrive_issuer = [&](int idx) {
    if (!k2 || cta_rank == 0u) mbar_arrive(bar(idx)); else mbar_arrive_remote_lite(bar(idx), 0u);
  };

  // ---- aggregation epilogue, sum phase: task = (sample of the tile, 8 of the pass's 128 columns); shared by the epilogue threads
  //      and, in inference pair mode, the input warps (u = thread number among the n_threads that take tasks) ----
  auto agg_one = [&](int task, int pass, int buf, int s_begin) {
    const float* stage = reinterpret_cast<const float*>(sA + 4 * kTileBytesA);
    const uint8_t* samp_row = misc + kOffSampRow + buf * 128;
    const uint8_t* samp_cnt = misc + kOffSampCnt + buf * 128;
    {
      const int sl = task >> 4, c8 = task & 15;
      const int r0 = samp_row[sl], cnt = samp_cnt[sl];
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      // a sample has 1..8 rows: two groups of four predicated, independent loads (same summation order as a plain loop)
#pragma unroll
      for (int j0 = 0; j0 < kK; j0 += 4) {
        if (j0 < cnt) {
          float4 t0[4], t1[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int r = min(r0 + j0 + j, 127);
            const float* rp = stage + r * 128 + ((c8 ^ (r & 7)) << 2);
            t0[j] = *reinterpret_cast<const float4*>(rp);
            t1[j] = *reinterpret_cast<const float4*>(rp + 64);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j0 + j < cnt) {
              acc[0] += t0[j].x; acc[1] += t0[j].y; acc[2] += t0[j].z; acc[3] += t0[j].w;
              acc[4] += t1[j].x; acc[5] += t1[j].y; acc[6] += t1[j].z; acc[7] += t1[j].w;
            }
          }
        }
      }
      const long long s = (long long)s_begin + sl;
      const int col = 128 * pass + 8 * c8;
      uint8_t* kbp = P.img + (size_t)(s >> 7) * kImgTileBytes + (size_t)(col >> 6) * (2 * kTileBytesA);
      if (!kF8) {
        uint4 hi, lo;
        split8(acc, hi, lo);
        uint8_t* p = kbp + swz((int)(s & 127), (col & 63) >> 3);
        *reinterpret_cast<uint4*>(p) = hi;
        *reinterpret_cast<uint4*>(p + kTileBytesA) = lo;
      } else {  // acc already carries kF8ActScale (the layer-3 epilogue scale and bias include it)
        uint4 hi;
        uint2 lo8, hi8;
        split8_f8(acc, hi, lo8, hi8);
        *reinterpret_cast<uint4*>(kbp + swz((int)(s & 127), (col & 63) >> 3)) = hi;
        *reinterpret_cast<uint2*>(kbp + kTileBytesA + swz8_lo((int)(s & 127), (col & 63) >> 3)) = lo8;
        *reinterpret_cast<uint2*>(kbp + kTileBytesA + swz8_hi((int)(s & 127), (col & 63) >> 3)) = hi8;
      }
    }
  };
  auto agg_sum_tasks = [&](int u, int n_threads, int pass, int buf) {
    const int* info = reinterpret_cast<const int*>(misc + kOffInfo) + buf * 4;
    const int s_begin = info[0], n_samp = info[1];
    for (int task = u; task < (NPCD_EXP_NOAGG ? 0 : n_samp * 16); task += n_threads) agg_one(task, pass, buf, s_begin);
  };
  // barrier of the aggregation passes: the 256 epilogue threads (+ the 128 input-warp threads in inference pair mode)
  auto agg_bar = [&]() {
    if (kPro) asm volatile("bar.sync 3, 384;" ::: "memory"); else epi_bar_sync();
  };

  if (warp == 0) {
    // ================================================= weight producer ==================================================
    // (whole warp runs the loop so every value stays warp-uniform; one elected lane issues the copies)
    int st = 0;
    uint32_t ph = 0;
    for (int pass = 0; pass < n_pass; ++pass) {
      for (int l = 0; l < n_layers; ++l) {
        const int nkb = (P.layers[l].ksteps + 3) >> 2;
        const uint8_t* src = P.layers[l].w;
        for (int t = 0; t < 2 * nkb; ++t) {
          mbar_wait(bar(kWE + st), ph ^ 1);  // cluster: both CTAs have drained this stage
          if (NPCD_EXP_NOREFILL) {
            if (elect_one()) mbar_arrive(bar(kWF + st));
          } else if (elect_one()) {
            mbar_expect_tx(bar(kWF + st), kStageBytes);
            if (k2)  // this CTA's half of the tile: output rows [128 rank, 128 rank + 128)
              bulk_g2s(smem_u32(sW + st * kStageBytes), src + (size_t)t * kTileBytesW + (size_t)cta_rank * kStageBytes, kStageBytes,
                       bar(kWF + st));
            else if (!kCluster)
              bulk_g2s(smem_u32(sW + st * kTileBytesW), src + (size_t)t * kTileBytesW, kTileBytesW, bar(kBarWFull + st));
            else if (cta_rank == 0)
              bulk_g2s_mc(smem_u32(sW + st * kTileBytesW), src + (size_t)t * kTileBytesW, kTileBytesW, bar(kBarWFull + st), 0x3);
          }
          __syncwarp();
          if (++st == kSt) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer ===================================================
    // The whole warp walks the schedule and waits on the barriers (all values warp-uniform -> uniform registers, no per-lane
    // descriptor shuffling); one elected lane -- always the same, tcgen05.commit tracks the issuing thread -- issues the MMAs.
    int st = 0;
    uint32_t ph_w = 0;
    uint32_t ph_ar = 0, ph_a0 = 0;  // one parity bit per K-block
    uint32_t ph_ar2 = 0;            // second half of K-block 0 (kSplit0)
    uint32_t ph_af = 0;             // acc_free parity bits (bit b = accumulator b)
    uint32_t lc = 0;                // running layer counter: layer lc accumulates into TMEM columns (lc & 1) * 256
    const uint64_t desc_a0 = make_desc(smem_u32(sA));
    const uint64_t desc_w0 = make_desc(smem_u32(sW));
    if (k2 && cta_rank != 0u) {
      // 2-SM, peer CTA: no MMAs are issued here; tell the leader when this CTA's half of every weight stage has landed
      for (int pass = 0; pass < n_pass; ++pass)
        for (int l = 0; l < n_layers; ++l) {
          const int nkb = (P.layers[l].ksteps + 3) >> 2;
          for (int t = 0; t < 2 * nkb; ++t) {
            mbar_wait(bar(kWF + st), ph_w);
            if (elect_one()) mbar_arrive_remote_lite(bar(kBar2WPeer + st), 0u);
            __syncwarp();
            if (++st == kSt) { st = 0; ph_w ^= 1; }
          }
        }
    } else
    for (int pass = 0, tile = blockIdx.x; pass < n_pass; ++pass, tile += gridDim.x) {
      if (!k2 && tile >= n_tiles) {
        // the partner CTA still has a tile: keep the shared weight ring turning (no MMAs here, plain arrives on both empty-barriers)
        for (int l = 0; l < n_layers; ++l) {
          const int nkb = (P.layers[l].ksteps + 3) >> 2;
          for (int t = 0; t < 2 * nkb; ++t) {
            mbar_wait(bar(kBarWFull + st), ph_w);
            if (elect_one()) {
              mbar_arrive(bar(kBarWEmpty + st));
              mbar_arrive_remote(bar(kBarWEmpty + st), cta_rank ^ 1u);
            }
            __syncwarp();
            if (++st == kStages) { st = 0; ph_w ^= 1; }
          }
        }
        continue;
      }
      // (2-SM: the waiters of both CTAs go on as long as the PAIR has another pass)
      const bool has_next = k2 ? pass + 1 < n_pass : tile + (int)gridDim.x < n_tiles;
      auto mma16 = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t accumulate) {
        if (NPCD_EXP_NSPLIT) {  // experiment: the same product as two N = 128 instructions (is the A re-read affordable?)
          umma_f16(d, a, b, make_idesc(128, 128), accumulate);
          umma_f16(d + 128u, a, b + 1024u, make_idesc(128, 128), accumulate);
          return;
        }
        if (k2) umma_f16_2sm(d, a, b, kIdescMma, accumulate); else umma_f16(d, a, b, kIdescMma, accumulate);
      };
      auto mma8 = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t accumulate) {
        if (NPCD_EXP_NOF8MMA) return;
        if (NPCD_EXP_NSPLIT) {
          umma_f8(d, a, b, make_idesc(128, 128), accumulate);
          umma_f8(d + 128u, a, b + 1024u, make_idesc(128, 128), accumulate);
          return;
        }
        if (k2) umma_f8_2sm(d, a, b, kIdescMma, accumulate); else umma_f8(d, a, b, kIdescMma, accumulate);
      };
      auto mma16ts = [&](uint32_t d, uint32_t a, uint64_t b, uint32_t accumulate) {  // A operand in tensor memory (kTS)
        if (k2) umma_f16_ts_2sm(d, a, b, kIdescMma, accumulate); else umma_f16_ts(d, a, b, kIdescMma, accumulate);
      };
      auto mma8ts = [&](uint32_t d, uint32_t a, uint64_t b, uint32_t accumulate) {
        if (k2) umma_f8_ts_2sm(d, a, b, kIdescMma, accumulate); else umma_f8_ts(d, a, b, kIdescMma, accumulate);
      };
      auto commit_stage = [&](int stage) {  // the weight stage may be refilled once the MMAs issued so far have completed
        if (k2) umma_commit_2sm(bar(kWE + stage));
        else if (kCluster) umma_commit_mc(bar(kBarWEmpty + stage), 0x3);
        else umma_commit(bar(kBarWEmpty + stage));
      };
      auto commit_cta = [&](int idx) {  // barriers the warps of each CTA wait on locally (2-SM: both CTAs' copies)
        if (k2) umma_commit_2sm(bar(idx)); else umma_commit(bar(idx));
      };
      long long tw_w = 0, tw_a = 0, tw_f = 0;  // timeline: cycles this tile's issue loop waited for weights / operands / accumulators
      auto timed_wait = [&](uint32_t b, uint32_t par, long long& acc) {
        if (P.timeline) {
          const long long t0 = clock64();
          mbar_wait(b, par);
          acc += clock64() - t0;
        } else {
          mbar_wait(b, par);
        }
      };
      for (int l = 0; l < n_layers; ++l, ++lc) {
        if (lane == 0 && l < 4) NPCD_TL(pass, 16 + 3 * l);
        const uint32_t ab = lc & 1u;
        const uint32_t d_tmem = tmem_base + ab * 256u;
        timed_wait(bar(kBarAccFree + ab), ((ph_af >> ab) & 1u) ^ 1u, tw_f);  // every epilogue warp has drained this accumulator
        ph_af ^= 1u << ab;
        const int ksteps = P.layers[l].ksteps;
        const int nkb = (ksteps + 3) >> 2;
        // heads layer 2 (channel_net.0) reads the same operand (feat) as layer 1 (shape_net.0): nothing new to wait for
        const bool fresh_a = !(kHeads && l + P.layer_ofs == 2);
        for (int kb = 0; kb < nkb; ++kb) {
          if ((!kPair || kPro) && l == 0) {  // first operand of a tile: the loader warp (heads) / the input warps (pair)
            timed_wait(bar(kBarA0Rdy + kb), (ph_a0 >> kb) & 1u, tw_a);
            if (k2 && !kPro) timed_wait(bar(kBar2A0Peer + kb), (ph_a0 >> kb) & 1u, tw_a);  // the peer's loader relays its copy
            ph_a0 ^= 1u << kb;
          } else if (fresh_a) {
            timed_wait(bar(kBarARdy + kb), (ph_ar >> kb) & 1u, tw_a);
            ph_ar ^= 1u << kb;
          }
          if (lane == 0 && kb == 0 && l < 4) NPCD_TL(pass, 17 + 3 * l);  // first operand block of the layer is there
          const int ks_n = min(4, ksteps - kb * 4);
          // descriptor start-address field counts 16-byte units: +2 per 16-column K step, +1024 per 16 KB tile
          const uint64_t a_hi = desc_a0 + (uint64_t)(kb * 2 * (kTileBytesA >> 4)), a_lo = a_hi + (kTileBytesA >> 4);
          if (kSplit0 && kb == 0 && fresh_a && l > 0) {
            // K-block 0 of an epilogue-produced operand arrives in two halves: columns {0-15, 32-47} (K16 steps 0, 2 of the fp16
            // tile; the first K32 step of either half of the 8-bit tile), then {16-31, 48-63} (steps 1, 3; the second K32 steps).
            // Both weight stages of the block are held until the second half has been issued.
            const int st2 = st + 1 == kSt ? 0 : st + 1;
            const uint32_t ph_w2 = st + 1 == kSt ? ph_w ^ 1u : ph_w;
            timed_wait(bar(kWF + st), ph_w, tw_w);
            timed_wait(bar(kWF + st2), ph_w2, tw_w);
            if (k2) {
              timed_wait(bar(kBar2WPeer + st), ph_w, tw_w);
              timed_wait(bar(kBar2WPeer + st2), ph_w2, tw_w);
            }
            tc_fence_after();
            const uint64_t b = desc_w0 + (uint64_t)(st * (kStageBytes >> 4)), b2 = desc_w0 + (uint64_t)(st2 * (kStageBytes >> 4));
            if (kTS && l >= kTsFirst) {
              // A from tensor memory: feature group t = 2 h + p of the block (thread half h, 16-value piece p) sits in columns
              // [32 h + 16 p, + 16) of the OTHER accumulator buffer: fp16 in the first 8, [lo8 | hi8] in the last 8
              const uint32_t a_t = tmem_base + (ab ^ 1u) * 256u;
              if (elect_one()) {
                mma16ts(d_tmem, a_t, b, 0u);
                mma16ts(d_tmem, a_t + 32u, b + 4, 1u);
                mma8ts(d_tmem, a_t + 8u, b2, 1u);
                mma8ts(d_tmem, a_t + 40u, b2 + 4, 1u);
              }
              __syncwarp();
              timed_wait(bar(kBarARdy2), ph_ar2, tw_a);
              ph_ar2 ^= 1u;
              tc_fence_after();
              if (elect_one()) {
                mma16ts(d_tmem, a_t + 16u, b + 2, 1u);
                mma16ts(d_tmem, a_t + 48u, b + 6, 1u);
                commit_stage(st);
                mma8ts(d_tmem, a_t + 24u, b2 + 2, 1u);
                mma8ts(d_tmem, a_t + 56u, b2 + 6, 1u);
                commit_stage(st2);
              }
              __syncwarp();
              for (int q = 0; q < 2; ++q)
                if (++st == kSt) { st = 0; ph_w ^= 1; }
              continue;
            }
            if (elect_one()) {
              mma16(d_tmem, a_hi, b, 0u);
              mma16(d_tmem, a_hi + 4, b + 4, 1u);
              if (!kF8) {
                mma16(d_tmem, a_lo, b, 1u);
                mma16(d_tmem, a_lo + 4, b + 4, 1u);
                mma16(d_tmem, a_hi, b2, 1u);
                mma16(d_tmem, a_hi + 4, b2 + 4, 1u);
              } else {
                mma8(d_tmem, a_lo, b2, 1u);
                if (!NPCD_EXP_NOWCORR && !P.no_wcorr) mma8(d_tmem, a_lo + 4, b2 + 4, 1u);
              }
            }
            __syncwarp();
            timed_wait(bar(kBarARdy2), ph_ar2, tw_a);
            ph_ar2 ^= 1u;
            tc_fence_after();
            if (elect_one()) {
              mma16(d_tmem, a_hi + 2, b + 2, 1u);
              mma16(d_tmem, a_hi + 6, b + 6, 1u);
              if (!kF8) {
                mma16(d_tmem, a_lo + 2, b + 2, 1u);
                mma16(d_tmem, a_lo + 6, b + 6, 1u);
              }
              commit_stage(st);
              if (!kF8) {
                mma16(d_tmem, a_hi + 2, b2 + 2, 1u);
                mma16(d_tmem, a_hi + 6, b2 + 6, 1u);
              } else {
                mma8(d_tmem, a_lo + 2, b2 + 2, 1u);
                if (!NPCD_EXP_NOWCORR && !P.no_wcorr) mma8(d_tmem, a_lo + 6, b2 + 6, 1u);
              }
              commit_stage(st2);
              if (l == n_layers - 1 && has_next && (!kPair || kb < 2)) commit_cta(kBarAFree + kb);
            }
            __syncwarp();
            for (int q = 0; q < 2; ++q)
              if (++st == kSt) { st = 0; ph_w ^= 1; }
            continue;
          }
          if (kTS && l >= kTsFirst) {  // K-blocks 1..3 of a tensor-memory operand: four K16 steps on the fp16 stage, four K32 steps on the 8-bit stage
            const uint32_t a_t = tmem_base + (ab ^ 1u) * 256u + 64u * (uint32_t)kb;
            timed_wait(bar(kWF + st), ph_w, tw_w);
            if (k2) timed_wait(bar(kBar2WPeer + st), ph_w, tw_w);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t b = desc_w0 + (uint64_t)(st * (kStageBytes >> 4));
#pragma unroll
              for (int t = 0; t < 4; ++t) mma16ts(d_tmem, a_t + 16u * t, b + 2 * t, 1u);
              commit_stage(st);
            }
            __syncwarp();
            if (++st == kSt) { st = 0; ph_w ^= 1; }
            timed_wait(bar(kWF + st), ph_w, tw_w);
            if (k2) timed_wait(bar(kBar2WPeer + st), ph_w, tw_w);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t b = desc_w0 + (uint64_t)(st * (kStageBytes >> 4));
#pragma unroll
              for (int t = 0; t < 4; ++t) mma8ts(d_tmem, a_t + 16u * t + 8u, b + 2 * t, 1u);
              commit_stage(st);
              if (kb == nkb - 1) commit_cta(kBarAccRdy + ab);
            }
            if (lane == 0 && kb == nkb - 1 && l < 4) NPCD_TL(pass, 18 + 3 * l);
            __syncwarp();
            if (++st == kSt) { st = 0; ph_w ^= 1; }
            continue;
          }
          // stage "hi": A_hi*W_hi + A_lo*W_hi
          timed_wait(bar(kWF + st), ph_w, tw_w);
          if (k2) timed_wait(bar(kBar2WPeer + st), ph_w, tw_w);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t b = desc_w0 + (uint64_t)(st * (kStageBytes >> 4));
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              if (ks < ks_n) mma16(d_tmem, a_hi + 2 * ks, b + 2 * ks, (kb | ks) != 0);
            if (!kF8) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                if (ks < ks_n) mma16(d_tmem, a_lo + 2 * ks, b + 2 * ks, 1u);
            }
            commit_stage(st);
          }
          __syncwarp();
          if (++st == kSt) { st = 0; ph_w ^= 1; }
          // stage "lo": A_hi*W_lo
          timed_wait(bar(kWF + st), ph_w, tw_w);
          if (k2) timed_wait(bar(kBar2WPeer + st), ph_w, tw_w);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t b = desc_w0 + (uint64_t)(st * (kStageBytes >> 4));
            if (!kF8) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                if (ks < ks_n) mma16(d_tmem, a_hi + 2 * ks, b + 2 * ks, 1u);
            } else {
              // the 8-bit tile: row = [lo8 of the 64 columns | hi8 of the 64 columns] against [Whi8 | Wlo8]; a K = 32 step
              // covers two fp16 K-steps, so a K-block with ks_n fp16 steps takes ks_n / 2 steps in each 64-byte half
#pragma unroll
              for (int k8 = 0; k8 < 2; ++k8)
                if (2 * k8 < ks_n) mma8(d_tmem, a_lo + 2 * k8, b + 2 * k8, 1u);
#pragma unroll
              for (int k8 = 0; k8 < 2; ++k8)
                if (2 * k8 < ks_n && !NPCD_EXP_NOWCORR && !P.no_wcorr) mma8(d_tmem, a_lo + 4 + 2 * k8, b + 4 + 2 * k8, 1u);
            }
            commit_stage(st);
            // this K-block may take the next tile's first operand (arrive only where somebody waits: pair mode restages
            // K-blocks 0..1, and the last tile of a CTA has no successor)
            // (kTS: the last layer that reads its A operand from shared memory releases the K-blocks for the next tile's input)
            if (l == (kTS ? kTsFirst - 1 : n_layers - 1) && has_next && (!kPair || kb < 2)) commit_cta(kBarAFree + kb);
            if (kb == nkb - 1) commit_cta(kBarAccRdy + ab);
          }
          if (lane == 0 && kb == nkb - 1 && l < 4) NPCD_TL(pass, 18 + 3 * l);  // last MMA of the layer issued
          __syncwarp();
          if (++st == kSt) { st = 0; ph_w ^= 1; }
        }
      }
      if (lane == 0) { NPCD_TLV(pass, 28, tw_w); NPCD_TLV(pass, 29, tw_a); NPCD_TLV(pass, 30, tw_f); }
    }
  } else if (warp == 10) {
    // ======================================== heads / probe: first-operand loader =======================================
    if (kTrainP) {
      // ---- training stash: every published K-block of the A operand (X_0 from the prologue, X_1..X_3 from the epilogues of
      //      layers 0..2) is bulk-copied to HBM as-is (it already IS the operand image the backward GEMMs consume); the epilogue
      //      threads wait on kBarStash + kb before they overwrite that K-block.  Same event order as the epilogue code below.
      uint32_t ph_ar = 0;
      auto stash_block = [&](int kb, uint8_t* dst) {
        mbar_wait(bar(kBarARdy + kb), (ph_ar >> kb) & 1u);
        ph_ar ^= 1u << kb;
        if (elect_one()) {
          bulk_s2g(dst, smem_u32(sA + kb * 2 * kTileBytesA), 2 * kTileBytesA);
          bulk_commit();
          bulk_wait_read0();
          mbar_arrive(bar(kBarStash + kb));
        }
        __syncwarp();
      };
      int cur = -1;
      for (int it = -1;; ++it) {
        const bool prime = it < 0;
        const int target = prime ? (int)blockIdx.x : cur + (int)gridDim.x;
        const bool has_target = target < n_tiles;
        if (prime && !has_target) break;
        if (!prime) {
          for (int l = 0; l < 3; ++l)
            for (int kb = 0; kb < 4; ++kb) stash_block(kb, P.stash_x[l + 1] + ((size_t)cur * 4 + kb) * (2 * kTileBytesA));
        }
        if (has_target)
          for (int kb = 0; kb < 2; ++kb) stash_block(kb, P.stash_x[0] + ((size_t)target * 2 + kb) * (2 * kTileBytesA));
        if (!has_target) break;
        cur = target;
      }
      if (elect_one()) bulk_wait_all0();  // the copies must have landed before the CTA retires its shared memory
      __syncwarp();
    } else if (!kPair) {
      uint32_t it = 0;
      const int n_it = k2 ? n_pass : (n_tiles > (int)blockIdx.x ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0);
      for (int tile = blockIdx.x; (int)it < n_it; tile += gridDim.x, ++it) {
        for (int kb = 0; kb < 4; ++kb) {
          if (it > 0) mbar_wait(bar(kBarAFree + kb), (it - 1) & 1u);
          if (elect_one()) {
            if (tile < n_tiles) {
              mbar_expect_tx(bar(kBarA0Rdy + kb), 2 * kTileBytesA);
              bulk_g2s(smem_u32(sA + kb * 2 * kTileBytesA), P.img + (size_t)tile * kImgTileBytes + (size_t)kb * 2 * kTileBytesA,
                       2 * kTileBytesA, bar(kBarA0Rdy + kb));
            } else {
              mbar_arrive(bar(kBarA0Rdy + kb));  // phantom tile of the 2-SM pair: whatever the K-block holds is multiplied, nothing is stored
            }
          }
          __syncwarp();
        }
        if (k2 && cta_rank != 0u) {  // the leader's issuer must know that THIS CTA's operand has landed too
          for (int kb = 0; kb < 4; ++kb) {
            mbar_wait(bar(kBarA0Rdy + kb), it & 1u);
            if (elect_one()) mbar_arrive_remote_lite(bar(kBar2A0Peer + kb), 0u);
            __syncwarp();
          }
        }
      }
    }
  } else if (kPro && warp >= 11) {
    // ============================ inference pair mode: input warps (one thread per row of the tile) ====================
    // layer-0 input, 96 columns (see the column map at the epilogue threads' prologue below; the two code paths build the same
    // operand): chunks of 8 columns, K-block 0 = chunks 0..7, K-block 1 = chunks 8..11.
    const int pt = threadIdx.x - 352;  // 0..127 = tile row
    const int row = pt, x7 = row & 7;
    const uint32_t rowbase = (uint32_t)((row >> 3) * 1024 + x7 * 128);
    float* wts_all = reinterpret_cast<float*>(misc + kOffWts);
    uint8_t* row_samp_all = misc + kOffRowSamp;
    uint8_t* samp_row_all = misc + kOffSampRow;
    uint8_t* samp_cnt_all = misc + kOffSampCnt;
    int* info_all = reinterpret_cast<int*>(misc + kOffInfo);
    uint4 c_hi[8];
    uint4 c_lo[8];  // f8 scheme: {lo8 (8 B), hi8 (8 B)} of the chunk
    auto split_c = [&](const float (&y)[8], int c) {
      if (!kF8) {
        split8(y, c_hi[c], c_lo[c]);
      } else {
        const float ys[8] = {y[0] * kF8ActScale, y[1] * kF8ActScale, y[2] * kF8ActScale, y[3] * kF8ActScale,
                             y[4] * kF8ActScale, y[5] * kF8ActScale, y[6] * kF8ActScale, y[7] * kF8ActScale};
        uint2 lo8, hi8;
        split8_f8(ys, c_hi[c], lo8, hi8);
        c_lo[c] = make_uint4(lo8.x, lo8.y, hi8.x, hi8.y);
      }
    };
    auto store_c = [&](int kb, int c, int src) {  // chunk c (8 columns) of K-block kb
      uint8_t* t = sA + kb * 2 * kTileBytesA + rowbase;
      *reinterpret_cast<uint4*>(t + ((c ^ x7) << 4)) = c_hi[src];
      if (!kF8) {
        *reinterpret_cast<uint4*>(t + kTileBytesA + ((c ^ x7) << 4)) = c_lo[src];
      } else {
        // (c, c + 1), c even, share the 16-byte chunks of their lo8 / hi8 bytes: one 16-byte store each when the odd one comes
        if (c & 1) {
          *reinterpret_cast<uint4*>(t + kTileBytesA + ((f8_chunk(c, kb == 0) ^ x7) << 4)) =
              make_uint4(c_lo[src - 1].x, c_lo[src - 1].y, c_lo[src].x, c_lo[src].y);
          *reinterpret_cast<uint4*>(t + kTileBytesA + (((f8_chunk(c, kb == 0) + 4) ^ x7) << 4)) =
              make_uint4(c_lo[src - 1].z, c_lo[src - 1].w, c_lo[src].z, c_lo[src].w);
        }
      }
    };
    auto publish_a0 = [&](int kb) {
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) arrive_issuer(kBarA0Rdy + kb);
    };
    // (sin, cos) of d * 2^i * pi for n consecutive octaves starting with frequency fr0, interleaved into v[0 .. 2n)
    auto octaves = [&](float d, float fr0, int n, float* v) {
      float fr = fr0;
#pragma unroll
      for (int i = 0; i < kFreqs; ++i) {
        if (i < n) {
          float sn, cs;
          if (NPCD_EXP_NOPOSENC) { sn = d * fr; cs = fr; } else sincos_small(d * fr, sn, cs);
          v[2 * i] = sn;
          v[2 * i + 1] = cs;
          fr *= 2.0f;
        }
      }
    };
    // yield to a running layer epilogue (kYield): called between the work units of prepare().  Cannot deadlock: the flag is only
    // set inside epilogue_store, which never waits for the input warps while they are inside prepare() (the one barrier it shares
    // with them, the closing barrier of the aggregation's last sum phase, is passed by the input warps BEFORE they enter prepare).
    auto yield_to_epilogue = [&]() {
      if (kYield) {
        while (*reinterpret_cast<volatile int*>(misc + kOffEpiBusy)) __nanosleep(100);
      }
    };
    const float kPi = 3.14159274101257324f;
    // input of tile `tile` into K-blocks 0 / 1; wait_free: the K-blocks still hold the previous tile's X_3 until layer 3's MMAs pass
    auto prepare = [&](int tile, int buf, bool wait_free, uint32_t parity) {
      const bool phantom = tile >= n_tiles;  // 2-SM pair with an odd tile count: an all-padding tile
      const int s_begin = phantom ? 0 : __ldg(P.tile_start + tile), s_end = phantom ? 0 : __ldg(P.tile_start + tile + 1);
      const int n_samp = s_end - s_begin;
      const int base = __ldg(P.pair_off + s_begin);
      uint8_t* row_samp = row_samp_all + buf * 128;
      uint8_t* samp_row = samp_row_all + buf * 128;
      uint8_t* samp_cnt = samp_cnt_all + buf * 128;
      if (pt < n_samp) {
        const int o = __ldg(P.pair_off + s_begin + pt) - base;
        const int c = __ldg(P.pair_off + s_begin + pt + 1) - base - o;
        samp_row[pt] = (uint8_t)o;
        samp_cnt[pt] = (uint8_t)c;
        for (int j = 0; j < c; ++j) row_samp[o + j] = (uint8_t)pt;
      }
      if (pt == 0) {
        info_all[buf * 4 + 0] = s_begin;
        info_all[buf * 4 + 1] = n_samp;
        info_all[buf * 4 + 2] = __ldg(P.pair_off + s_end) - base;
      }
      asm volatile("bar.sync 2, 128;" ::: "memory");
      const int n_rows = info_all[buf * 4 + 2];
      int idx = -1;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      float px = 0.f, py = 0.f, pz = 0.f;
      if (row < n_rows) {
        const int sl = row_samp[row];
        const long long s = (long long)s_begin + sl;
        idx = __ldg(P.nbr_idx + s * kK + (row - samp_row[sl]));
        x = __ldg(P.sample_pos + s);
        px = __ldg(P.kp_pos + (size_t)idx * 3); py = __ldg(P.kp_pos + (size_t)idx * 3 + 1); pz = __ldg(P.kp_pos + (size_t)idx * 3 + 2);
      }
      const bool live = idx >= 0;
      const float d3[3] = {x.x - px, x.y - py, x.z - pz};
      const float nrm = sqrtf(d3[0] * d3[0] + d3[1] * d3[1] + d3[2] * d3[2]);
      wts_all[buf * 128 + row] = live ? 1.0f / (nrm + 1e-5f) : 0.f;
      // ---- K-block 0: [feat 0..31 | x: (sin, cos) of octaves 0..7 | d_x, (sin, cos)_x of octaves 8, 9, d_y, (sin, cos)_y of octaves 0..4]
      float4 ft[8];  // the point's 32 features: loads in flight while this warp yields
#pragma unroll
      for (int j = 0; j < 8; ++j)
        ft[j] = live ? __ldg(reinterpret_cast<const float4*>(P.kp_feat + (size_t)idx * 32 + j * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
      yield_to_epilogue();
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const float y[8] = {ft[2 * ch].x, ft[2 * ch].y, ft[2 * ch].z, ft[2 * ch].w, ft[2 * ch + 1].x, ft[2 * ch + 1].y, ft[2 * ch + 1].z,
                            ft[2 * ch + 1].w};
        split_c(y, ch);
      }
      yield_to_epilogue();
      {
        float v[32];
        octaves(d3[0], kPi, 8, v);
        v[16] = d3[0];
        octaves(d3[0], kPi * 256.0f, 2, v + 17);
        v[21] = d3[1];
        octaves(d3[1], kPi, 5, v + 22);
        if (!live) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const float y[8] = {v[ch * 8], v[ch * 8 + 1], v[ch * 8 + 2], v[ch * 8 + 3], v[ch * 8 + 4], v[ch * 8 + 5], v[ch * 8 + 6], v[ch * 8 + 7]};
          split_c(y, 4 + ch);
        }
      }
      if (wait_free) mbar_wait(bar(kBarAFree + 0), parity);
#pragma unroll
      for (int c = 0; c < 8; ++c) store_c(0, c, c);
      publish_a0(0);
      // ---- K-block 1: [(sin, cos)_y of octaves 5..9 | z: d, (sin, cos) x 10 | 0]
      yield_to_epilogue();
      {
        float v[32];
        octaves(d3[1], kPi * 32.0f, 5, v);
        v[10] = d3[2];
        octaves(d3[2], kPi, kFreqs, v + 11);
        v[31] = 0.f;
        if (!live) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const float y[8] = {v[ch * 8], v[ch * 8 + 1], v[ch * 8 + 2], v[ch * 8 + 3], v[ch * 8 + 4], v[ch * 8 + 5], v[ch * 8 + 6], v[ch * 8 + 7]};
          split_c(y, ch);
        }
      }
      if (wait_free) mbar_wait(bar(kBarAFree + 1), parity);
#pragma unroll
      for (int c = 0; c < 4; ++c) store_c(1, c, c);
      publish_a0(1);
    };
    const int n_it = k2 ? n_pass : (n_tiles > (int)blockIdx.x ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0);
    if (n_it > 0) prepare((int)blockIdx.x, 0, false, 0u);
    uint32_t it = 0;
    for (int tile = blockIdx.x; (int)it < n_it; tile += gridDim.x, ++it) {
      const int next = tile + (int)gridDim.x;
      if ((int)it + 1 < n_it) prepare(next, (int)((it + 1) & 1u), true, it & 1u);
      // a third of the segmented-sum tasks of this tile's aggregation epilogue (two passes of 128 columns)
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        agg_bar();  // the epilogue threads have staged the pass
        if (kDynAgg && pass == 1) agg_sum_tasks(pt, kProThreads, pass, (int)(it & 1u));  // alone: the epilogue warps are on the next tile
        else agg_sum_tasks(kEpiThreads + pt, kEpiThreads + kProThreads, pass, (int)(it & 1u));
        agg_bar();
      }
    }
  } else if (warp == 11) {
    // ================================ heads training: stash of F, C1, C2, C3 (operand images) ===========================
    if (kTrainH) {
      uint32_t ph_ar = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int j = 0; j < 4; ++j) {
          for (int kb = 0; kb < 4; ++kb) {
            mbar_wait(bar(kBarARdy + kb), (ph_ar >> kb) & 1u);
            ph_ar ^= 1u << kb;
            if (elect_one()) {
              bulk_s2g(P.hstash_x[j] + ((size_t)tile * 4 + kb) * (2 * kTileBytesA), smem_u32(sA + kb * 2 * kTileBytesA), 2 * kTileBytesA);
              bulk_commit();
              bulk_wait_read0();
              mbar_arrive(bar(kBarStash + kb));
              // C3 is out: the loader may take this K-block after the last MMAs (no successor tile: nobody waits)
              if (j == 3 && tile + (int)gridDim.x < n_tiles) mbar_arrive(bar(kBarAFree + kb));
            }
            __syncwarp();
          }
        }
      }
      if (elect_one()) bulk_wait_all0();
      __syncwarp();
    }
  } else {
    // ============================================ prologue / epilogue threads ===========================================
    const int et = threadIdx.x - 64;       // 0..255
    const int q = warp & 3;                // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;      // 0 / 1: which 32-column chunk of every 64-column K-block this thread handles
    const int row = q * 32 + lane;         // tile row == TMEM lane
    const int x7 = row & 7;
    const uint32_t rowbase = (uint32_t)((row >> 3) * 1024 + x7 * 128);
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t ph_acc = 0;                   // acc_ready parity bits
    uint32_t lc = 0;

    float* wts_all = reinterpret_cast<float*>(misc + kOffWts);
    uint8_t* row_samp_all = misc + kOffRowSamp;
    uint8_t* samp_row_all = misc + kOffSampRow;
    uint8_t* samp_cnt_all = misc + kOffSampCnt;
    int* info_all = reinterpret_cast<int*>(misc + kOffInfo);

    // publish one K-block of the A operand (all 8 epilogue warps arrive once per K-block)
    auto publish = [&](int barrier_index) {
      if (kTS) tmem_wait_st();  // (the layer-0 input of a kTS kernel is published by the input warps, not here)
      tc_fence_before();
      if (!kTS) fence_proxy_async();
      __syncwarp();
      if (lane == 0) arrive_issuer(barrier_index);
    };
    auto wait_acc = [&](uint32_t ab) {
      mbar_wait(bar(kBarAccRdy + ab), (ph_acc >> ab) & 1u);
      ph_acc ^= 1u << ab;
      tc_fence_after();
    };
    auto release_acc = [&](uint32_t ab) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_issuer(kBarAccFree + ab);
    };

    // training: K-block kb may only be overwritten once the stash warp has copied its previous contents out
    uint32_t sd_pending = 0, ph_sd = 0;
    int tile_now = 0;  // tile the layer epilogues currently work on (stash addressing)
    auto stash_wait = [&](int kb) {
      if (kTrain && ((sd_pending >> kb) & 1u)) {
        mbar_wait(bar(kBarStash + kb), (ph_sd >> kb) & 1u);
        ph_sd ^= 1u << kb;
        sd_pending &= ~(1u << kb);
      }
    };

    // ACT / LINEAR epilogue of layer l: four 32-column chunks (2 i + half), software-pipelined TMEM loads
    int tl_it = 0;  // tile counter of this CTA (timeline)
    // mid (inference pair kernel, layer 0 of every tile but a CTA's first): the epilogue starts while the input warps still sum
    // pass 1 of the PREVIOUS tile's aggregation out of K-blocks 2..3; K-blocks 0..1 are written first, then the closing barrier of
    // that sum phase, then K-blocks 2..3 (see the tile loop below).  A run-time flag on purpose: a second instantiation of this
    // lambda (+19 KB of code, cold whenever it was entered) started ~700 cycles later than the shared one.
    auto epilogue_store = [&](bool mid, int l, float slope, float* feat_row) {
      const uint32_t ab = lc & 1u;
      if (et == 0) NPCD_TL(tl_it, 2 * l);
      wait_acc(ab);
      if (kYield && et == 0) *reinterpret_cast<volatile int*>(misc + kOffEpiBusy) = 1;
      if (et == 0) NPCD_TL(tl_it, 2 * l + 1);
      const float inv = P.layers[l].inv_scale;
      const uint32_t t_acc = t_row + ab * 256u;
      uint32_t v[2][32];
      constexpr bool kSplitLd = kSplit0 && !NPCD_EXP_NOSPLITLD;
      if (kSplitLd) {  // first chunk in two 16-column loads, the second one issued AFTER the first half is published (tmem_ld16_async)
        tmem_ld16_async(t_acc + half * 32, &v[0][0]);
        tmem_wait16(&v[0][0]);
      } else {
        tmem_ld32_async(t_acc + half * 32, v[0]);
        tmem_wait(v[0]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (i < 3 && !(kSplitLd && i == 0)) tmem_ld32_async(t_acc + (2 * (i + 1) + half) * 32, v[(i + 1) & 1]);
        stash_wait(i);
        if (kDynAgg && i == 2 && mid) {
          agg_bar();  // pass 1 of the previous tile's aggregation has been summed: its staging area (K-blocks 2..3) is free
          if (et == 0) NPCD_TL(tl_it - 1, 15);
        }
        uint32_t* mptr = nullptr;
        if (kTrainP) mptr = P.stash_mask[l] + ((size_t)tile_now * 128 + row) * 8 + (2 * i + half);
        if (kTrainH && l >= 2) mptr = P.hstash_mask[l - 1] + ((size_t)tile_now * 128 + row) * 8 + (2 * i + half);
        if (kSplit0 && i == 0) {
          if (kTS) epi_chunk_ts<0, 2>(P, l, inv, slope, v[0], half * 32, t_acc + half * 32);
          else epi_chunk_store<kF8, 0, 2>(P, l, inv, slope, v[0], half * 32, sA, rowbase, x7, feat_row, nullptr);
          publish(kBarARdy + 0);
          if (kSplitLd) {
            tmem_ld16_async(t_acc + half * 32 + 16, &v[0][16]);
            tmem_ld32_async(t_acc + (2 + half) * 32, v[1]);
            tmem_wait16(&v[0][16]);
          }
          if (kTS) epi_chunk_ts<2, 4>(P, l, inv, slope, v[0], half * 32, t_acc + half * 32);
          else epi_chunk_store<kF8, 2, 4>(P, l, inv, slope, v[0], half * 32, sA, rowbase, x7, feat_row, nullptr);
          publish(kBarARdy2);
        } else {
          if (kTS) epi_chunk_ts<0, 4>(P, l, inv, slope, v[i & 1], (2 * i + half) * 32, t_acc + (2 * i + half) * 32);
          else epi_chunk_store<kF8>(P, l, inv, slope, v[i & 1], (2 * i + half) * 32, sA, rowbase, x7, feat_row, mptr);
          if (i == 3) release_acc(ab);
          publish(kBarARdy + i);
        }
        if (kTrain) sd_pending |= 1u << i;
        if (i < 3) tmem_wait(v[(i + 1) & 1]);
      }
      if (kYield && et == 0) *reinterpret_cast<volatile int*>(misc + kOffEpiBusy) = 0;
      ++lc;
    };
    constexpr bool kWhole = false;

    if (kPair) {
      // ------------------------------------------------------------------------------------------------- pair mode ----
      // layer-0 input, 96 columns = 6 K-steps for the 95 real ones (a 112-column layout with per-axis groups cost a seventh):
      //   [feat 0..31 | x: (sin, cos) of octaves 0..7 | d_x, (sin, cos)_x of octaves 8, 9 | y: d, (sin, cos) x 10 | z: d, (sin, cos) x 10 | 0]
      // (the column order is OURS; the first-layer weights are permuted to match when they are packed, ops.pair_input_perm).
      // half 0 owns columns 0..47 (K-block 0, chunks 0..5); half 1 owns 48..95 (K-block 0 chunks 6, 7; K-block 1 chunks 0..3).
      uint4 pre_hi[6], pre_lo[6];  // f8 scheme: pre_lo[c] = {lo8 (8 B), hi8 (8 B)} of the chunk

      auto split_pre = [&](const float (&y)[8], int c) {
        if (!kF8) {
          split8(y, pre_hi[c], pre_lo[c]);
        } else {
          const float ys[8] = {y[0] * kF8ActScale, y[1] * kF8ActScale, y[2] * kF8ActScale, y[3] * kF8ActScale,
                               y[4] * kF8ActScale, y[5] * kF8ActScale, y[6] * kF8ActScale, y[7] * kF8ActScale};
          uint2 lo8, hi8;
          split8_f8(ys, pre_hi[c], lo8, hi8);
          pre_lo[c] = make_uint4(lo8.x, lo8.y, hi8.x, hi8.y);
        }
      };
      // store chunk c (8 columns) of K-block kb of the staged layer-0 input
      auto store_pre = [&](int kb, int c, int src) {
        uint8_t* t = sA + kb * 2 * kTileBytesA + rowbase;
        *reinterpret_cast<uint4*>(t + ((c ^ x7) << 4)) = pre_hi[src];
        if (!kF8) {
          *reinterpret_cast<uint4*>(t + kTileBytesA + ((c ^ x7) << 4)) = pre_lo[src];
        } else {
          *reinterpret_cast<uint2*>(t + kTileBytesA + ((f8_chunk(c, kb == 0) ^ x7) << 4) + (c & 1) * 8) = make_uint2(pre_lo[src].x, pre_lo[src].y);
          *reinterpret_cast<uint2*>(t + kTileBytesA + (((f8_chunk(c, kb == 0) + 4) ^ x7) << 4) + (c & 1) * 8) = make_uint2(pre_lo[src].z, pre_lo[src].w);
        }
      };

      auto prologue_compute = [&](int tile, int buf) {
        const int s_begin = __ldg(P.tile_start + tile), s_end = __ldg(P.tile_start + tile + 1);
        const int n_samp = s_end - s_begin;
        const int base = __ldg(P.pair_off + s_begin);
        uint8_t* row_samp = row_samp_all + buf * 128;
        uint8_t* samp_row = samp_row_all + buf * 128;
        uint8_t* samp_cnt = samp_cnt_all + buf * 128;
        if (et < n_samp) {
          const int o = __ldg(P.pair_off + s_begin + et) - base;
          const int c = __ldg(P.pair_off + s_begin + et + 1) - base - o;
          samp_row[et] = (uint8_t)o;
          samp_cnt[et] = (uint8_t)c;
          for (int j = 0; j < c; ++j) row_samp[o + j] = (uint8_t)et;
        }
        if (et == 0) {
          info_all[buf * 4 + 0] = s_begin;
          info_all[buf * 4 + 1] = n_samp;
          info_all[buf * 4 + 2] = __ldg(P.pair_off + s_end) - base;
        }
        epi_bar_sync();
        const int n_rows = info_all[buf * 4 + 2];
        int idx = -1;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        float px = 0.f, py = 0.f, pz = 0.f;
        if (row < n_rows) {
          const int sl = row_samp[row];
          const long long s = (long long)s_begin + sl;
          idx = __ldg(P.nbr_idx + s * kK + (row - samp_row[sl]));
          x = __ldg(P.sample_pos + s);
          px = __ldg(P.kp_pos + (size_t)idx * 3); py = __ldg(P.kp_pos + (size_t)idx * 3 + 1); pz = __ldg(P.kp_pos + (size_t)idx * 3 + 2);
        }
        if (kTrain && half == 0) {
          P.stash_idx[(size_t)tile * 128 + row] = idx;
          P.stash_samp[(size_t)tile * 128 + row] = row < n_rows ? s_begin + (int)row_samp[row] : -1;
        }
        const float d3[3] = {x.x - px, x.y - py, x.z - pz};
        // (sin, cos) of d * 2^i * pi for n consecutive octaves starting with frequency fr0, interleaved into v[0 .. 2n)
        auto octaves = [&](float d, float fr0, int n, float* v) {
          float fr = fr0;
#pragma unroll
          for (int i = 0; i < kFreqs; ++i) {
            if (i < n) {
              float sn, cs;
              sincos_small(d * fr, sn, cs);
              v[2 * i] = sn;
              v[2 * i + 1] = cs;
              fr *= 2.0f;
            }
          }
        };
        const float kPi = 3.14159274101257324f;
        if (half == 0) {
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            float y[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (idx >= 0) {
              const float4 f0 = __ldg(reinterpret_cast<const float4*>(P.kp_feat + (size_t)idx * 32 + ch * 8));
              const float4 f1 = __ldg(reinterpret_cast<const float4*>(P.kp_feat + (size_t)idx * 32 + ch * 8 + 4));
              y[0] = f0.x; y[1] = f0.y; y[2] = f0.z; y[3] = f0.w; y[4] = f1.x; y[5] = f1.y; y[6] = f1.z; y[7] = f1.w;
            }
            split_pre(y, ch);
          }
          float v[16];
          octaves(d3[0], kPi, 8, v);
          if (idx < 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = 0.f;
          }
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            const float y[8] = {v[ch * 8], v[ch * 8 + 1], v[ch * 8 + 2], v[ch * 8 + 3], v[ch * 8 + 4], v[ch * 8 + 5], v[ch * 8 + 6], v[ch * 8 + 7]};
            split_pre(y, 4 + ch);
          }
          const float nrm = sqrtf(d3[0] * d3[0] + d3[1] * d3[1] + d3[2] * d3[2]);
          wts_all[buf * 128 + row] = idx >= 0 ? 1.0f / (nrm + 1e-5f) : 0.f;
        } else {
          float v[48];
          v[0] = d3[0];
          octaves(d3[0], kPi * 256.0f, 2, v + 1);
          v[5] = d3[1];
          octaves(d3[1], kPi, kFreqs, v + 6);
          v[26] = d3[2];
          octaves(d3[2], kPi, kFreqs, v + 27);
          v[47] = 0.f;
          if (idx < 0) {
#pragma unroll
            for (int i = 0; i < 48; ++i) v[i] = 0.f;
          }
#pragma unroll
          for (int ch = 0; ch < 6; ++ch) {
            const float y[8] = {v[ch * 8], v[ch * 8 + 1], v[ch * 8 + 2], v[ch * 8 + 3], v[ch * 8 + 4], v[ch * 8 + 5], v[ch * 8 + 6], v[ch * 8 + 7]};
            split_pre(y, ch);
          }
        }
      };
      // store the K-block-kb part of the staged layer-0 input and publish that K-block
      auto prologue_store = [&](int kb) {
        stash_wait(kb);
        if (kb == 0) {
          if (half == 0) {
#pragma unroll
            for (int c = 0; c < 6; ++c) store_pre(0, c, c);
          } else {
#pragma unroll
            for (int c = 0; c < 2; ++c) store_pre(0, 6 + c, c);
          }
        } else if (half == 1) {
#pragma unroll
          for (int c = 0; c < 4; ++c) store_pre(1, c, c + 2);
        }
        publish(kBarARdy + kb);
        if (kTrain) sd_pending |= 1u << kb;
      };

      // ---- layer 3: bias + LeakyReLU, normalised inverse-distance weight, segmented sum over each sample's rows
      //      (fields/aggregators/mlp.py:86-88,119-121), staged as fp32 in K-blocks 2..3 (free once layer 3's MMAs are done),
      //      two passes of 128 columns; the sums leave as the pre-split operand image the heads kernel bulk-copies.
      //      defer_last (inference pair kernel): returns once pass 1 is staged and the accumulator released; the caller runs the
      //      last barrier / sum / barrier itself, with the first half of the next tile's layer-0 epilogue in between.
      auto agg_epilogue = [&](int buf, int cur_tile, bool defer_last) {
        const uint32_t ab = lc & 1u;
        if (et == 0) NPCD_TL(tl_it, 6);
        wait_acc(ab);
        if (et == 0) NPCD_TL(tl_it, 7);
        if (!kPro) epi_bar_sync();  // wts[] / row maps of this tile were written by other warps of this group (kPro: by the input
                                    // warps a tile ago, ordered by their A0Rdy arrive -> MMA -> accumulator-ready chain)
        const float inv = P.layers[3].inv_scale;
        const uint32_t t_acc = t_row + ab * 256u;
        const float* wts = wts_all + buf * 128;
        const uint8_t* row_samp = row_samp_all + buf * 128;
        const uint8_t* samp_row = samp_row_all + buf * 128;
        const uint8_t* samp_cnt = samp_cnt_all + buf * 128;
        const int n_rows = info_all[buf * 4 + 2];
        float wn = 0.f;
        if (row < n_rows) {
          const int sl = row_samp[row];
          const int r0 = samp_row[sl], cnt = samp_cnt[sl];
          float wsum = 0.f;
          for (int j = 0; j < cnt; ++j) wsum += wts[r0 + j];
          wn = wts[row] / wsum;
        }
        if (kTrain && half == 0) P.stash_wn[(size_t)cur_tile * 128 + row] = wn;
        stash_wait(2);  // the fp32 staging below overwrites K-blocks 2..3 (X_3)
        stash_wait(3);
        // staging: [128 rows][128 cols] fp32 in K-blocks 2..3.  16-byte chunk c4 of a row sits at position
        // ((c4 >> 1) | ((c4 & 1) << 4)) ^ (row & 7): conflict-free for the row-per-lane stores AND for the sum phase, where
        // 16 lanes read chunks 2 c8 and 2 c8 + 1 of one row.
        float* stage = reinterpret_cast<float*>(sA + 4 * kTileBytesA);
        const uint64_t inv2 = pack2(inv, inv), slope2 = pack2(0.01f, 0.01f), wn2 = pack2(wn, wn);
        // four 32-column chunks q (pass = q >> 1): the TMEM load of chunk q + 1 is in flight while chunk q is processed
        uint32_t v[2][32];
        tmem_ld32_async(t_acc + 64 * half, v[0]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int pass = q >> 1;
          const int cl = 64 * half + 32 * (q & 1);  // column within the pass
          const int c0 = 128 * pass + cl;
          tmem_wait(v[q & 1]);
          if (q < 3) tmem_ld32_async(t_acc + 128 * ((q + 1) >> 1) + 64 * half + 32 * ((q + 1) & 1), v[(q + 1) & 1]);
          uint32_t mbits = 0u;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 b = *reinterpret_cast<const float4*>(&P.bias[3][c0 + g * 4]);
            const uint32_t* vv = v[q & 1];
            float y0, y1, y2, y3;
            act2(vv[g * 4 + 0], vv[g * 4 + 1], inv2, b.x, b.y, slope2, y0, y1);
            act2(vv[g * 4 + 2], vv[g * 4 + 3], inv2, b.z, b.w, slope2, y2, y3);
            if (kTrain)
              mbits |= ((y0 > 0.f ? 1u : 0u) | (y1 > 0.f ? 2u : 0u) | (y2 > 0.f ? 4u : 0u) | (y3 > 0.f ? 8u : 0u)) << (g * 4);
            float4 o;
            unpack2(mul2(pack2(y0, y1), wn2), o.x, o.y);
            unpack2(mul2(pack2(y2, y3), wn2), o.z, o.w);
            const int c4 = (cl >> 2) + g;
            const int pos = ((c4 >> 1) | ((c4 & 1) << 4)) ^ x7;
            *reinterpret_cast<float4*>(stage + row * 128 + (pos << 2)) = o;
          }
          if (kTrain) P.stash_mask[3][((size_t)cur_tile * 128 + row) * 8 + (c0 >> 5)] = mbits;
          if (q & 1) {
            if (pass == 1) release_acc(ab);
            if (et == 0) NPCD_TL(tl_it, 8 + 4 * pass);
            if (!(pass == 1 && defer_last)) {
              agg_bar();
              if (et == 0) NPCD_TL(tl_it, 9 + 4 * pass);
              agg_sum_tasks(et, kPro ? kEpiThreads + kProThreads : kEpiThreads, pass, buf);
              if (et == 0) NPCD_TL(tl_it, 10 + 4 * pass);
              agg_bar();  // the staging area is rewritten by the next pass / the next tile's layer-0 epilogue
              if (et == 0) NPCD_TL(tl_it, 11 + 4 * pass);
            }
          }
        }
        ++tl_it;
        ++lc;
      };

      int cur = kPro ? (int)blockIdx.x : -1;
      if (kPro) tile_now = cur;
      if (kDynAgg) {
        // Inference pair kernel (1-SM).  The aggregation of tile t used to keep the epilogue warps busy for ~5 700 cycles during which
        // the tensor pipe only had layer 0 of tile t + 1 (timeline: profiles/r2_timeline_pair.md, "the largest single gap").  Now,
        // as soon as pass 1 is STAGED (accumulator released), the epilogue warps start tile t + 1's layer-0 epilogue (K-blocks 0..1
        // first; the staging area is K-blocks 2..3), so layer 1's MMAs start ~2 000 cycles earlier, while the input warps ALONE
        // sum pass 1; the closing barrier of that sum phase sits in front of K-block 2 of the layer-0 epilogue.
        // (one call site of epilogue_store: one epilogue per trip, the aggregation after every third)
        bool mid = false;
        int buf = 0;
#pragma unroll 1
        for (int l = 0; cur < n_tiles;) {
          epilogue_store(mid, l, 0.01f, nullptr);
          mid = false;
          if (++l < 3) continue;
          agg_epilogue(buf, cur, true);
          agg_bar();  // pass 1 is staged: the input warps sum it
          if (et == 0) NPCD_TL(tl_it - 1, 13);
          cur += (int)gridDim.x;
          if (cur >= n_tiles) {
            agg_bar();  // (closing barrier of the sum phase; with a next tile it sits inside that tile's layer-0 epilogue)
            break;
          }
          tile_now = cur;
          buf ^= 1;
          l = 0;
          mid = true;
        }
      } else
      // Iteration -1 primes the pipeline (stages the first tile's input); iteration `it` runs the four layer epilogues of tile
      // `cur` and, between them, stages the input of the tile after it.  One call site per helper keeps the code I-cache sized.
      // (kPro, inference: the input warps 11..14 stage the layer-0 input; no priming iteration, no prologue work here.)
      for (int it = kPro ? 0 : -1;; ++it) {
        const bool prime = it < 0;
        if (kPro && (k2 ? it >= n_pass : cur >= n_tiles)) break;  // (2-SM: phantom tiles keep the pair in lockstep)
        const int target = prime ? (int)blockIdx.x : cur + (int)gridDim.x;  // tile whose layer-0 input is staged now
        const bool has_target = target < n_tiles;
        if (prime && !has_target) break;
        const int buf = it & 1;
#pragma unroll 1
        for (int l = 0; l < 3; ++l) {
          if (!prime) epilogue_store(kWhole, l, 0.01f, nullptr);
          if (!kPro && l == 0 && has_target) prologue_compute(target, buf ^ 1);
        }
        if (!kPro && has_target) {  // layer 3's MMAs release K-blocks 0 and 1 as they pass them
#pragma unroll 1
          for (int kb = 0; kb < 2; ++kb) {
            if (!prime) mbar_wait(bar(kBarAFree + kb), (uint32_t)it & 1u);
            prologue_store(kb);
          }
        }
        if (prime) { cur = target; tile_now = cur; continue; }
        agg_epilogue(buf, cur, false);
        if (!has_target && !k2) break;
        cur = target;
        tile_now = cur;
      }
    } else {
      // ------------------------------------------------------------------------------------------- heads / probe mode ----
      float* part = reinterpret_cast<float*>(misc + kOffPart);  // [128][4]
      // EPI_DOT1 (shape_net.2 -> softplus(x-1)) / EPI_DOT3 (channel_net.8 -> sigmoid): dot products of the activated row with the
      // 1 or 3 output weight vectors; the two threads of a row combine through shared memory.
      auto epilogue_dot = [&](int l, int nout, float (&res)[3]) {
        const uint32_t ab = lc & 1u;
        if (et == 0) NPCD_TL(tl_it, 2 * l);
        wait_acc(ab);
        if (et == 0) NPCD_TL(tl_it, 2 * l + 1);
        const float inv = P.layers[l].inv_scale;
        const uint32_t t_acc = t_row + ab * 256u;
        float acc3[3] = {0.f, 0.f, 0.f};
        // inference: all 128 accumulator columns of the thread go to registers first and the accumulator is released at once -- the
        // next layer that accumulates into it (after the three-output epilogue: channel_net.0' of the NEXT tile) no longer waits for
        // the ~6 000 cycles of dot-product arithmetic (352 threads: 184 registers per thread are there)
        constexpr bool kEarly = kMode == MODE_HEADS && !NPCD_EXP_NOEARLYREL;
        uint32_t v[kEarly ? 4 : 2][32];
        if (kEarly) {
#pragma unroll
          for (int i = 0; i < 4; ++i) tmem_ld32_async(t_acc + (2 * i + half) * 32, v[i]);
#pragma unroll
          for (int i = 0; i < 4; ++i) tmem_wait(v[i]);
          release_acc(ab);
        } else {
          tmem_ld32_async(t_acc + half * 32, v[0]);
          tmem_wait(v[0]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (!kEarly && i < 3) tmem_ld32_async(t_acc + (2 * (i + 1) + half) * 32, v[(i + 1) & 1]);
          const int c0 = (2 * i + half) * 32;
          uint32_t mbits = 0u;
          float hp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 b = *reinterpret_cast<const float4*>(&P.bias[l][c0 + g * 4]);
            const uint32_t* vv = v[kEarly ? i : (i & 1)];
            const float h0 = lrelu(fmaf(__uint_as_float(vv[g * 4 + 0]), inv, b.x)), h1 = lrelu(fmaf(__uint_as_float(vv[g * 4 + 1]), inv, b.y)),
                        h2 = lrelu(fmaf(__uint_as_float(vv[g * 4 + 2]), inv, b.z)), h3 = lrelu(fmaf(__uint_as_float(vv[g * 4 + 3]), inv, b.w));
            if (kTrainH) {  // the activated row is an operand of the narrow output layer's weight gradient: stash it as an image
              mbits |= ((h0 > 0.f ? 1u : 0u) | (h1 > 0.f ? 2u : 0u) | (h2 > 0.f ? 4u : 0u) | (h3 > 0.f ? 8u : 0u)) << (g * 4);
              if (g & 1) {
                const float y[8] = {hp[0], hp[1], hp[2], hp[3], h0, h1, h2, h3};
                uint4 hi, lo;
                split8(y, hi, lo);
                uint8_t* p = P.hstash_x[nout == 1 ? 5 : 4] + ((size_t)tile_now * 4 + (c0 >> 6)) * (2 * kTileBytesA) +
                             swz(row, ((c0 & 63) >> 3) + (g >> 1));
                *reinterpret_cast<uint4*>(p) = hi;
                *reinterpret_cast<uint4*>(p + kTileBytesA) = lo;
              } else {
                hp[0] = h0; hp[1] = h1; hp[2] = h2; hp[3] = h3;
              }
            }
            if (nout == 1) {
              const float4 w = *reinterpret_cast<const float4*>(&P.shape_out_w[c0 + g * 4]);
              acc3[0] = fmaf(h0, w.x, fmaf(h1, w.y, fmaf(h2, w.z, fmaf(h3, w.w, acc3[0]))));
            } else {
#pragma unroll
              for (int o = 0; o < 3; ++o) {
                const float4 w = *reinterpret_cast<const float4*>(&P.chan_out_w[o][c0 + g * 4]);
                acc3[o] = fmaf(h0, w.x, fmaf(h1, w.y, fmaf(h2, w.z, fmaf(h3, w.w, acc3[o]))));
              }
            }
          }
          if (kTrainH) P.hstash_mask[nout == 1 ? 0 : 4][((size_t)tile_now * 128 + row) * 8 + (2 * i + half)] = mbits;
          if (!kEarly && i < 3) tmem_wait(v[(i + 1) & 1]);
        }
        if (!kEarly) release_acc(ab);
        if (half == 1) { part[row * 4] = acc3[0]; part[row * 4 + 1] = acc3[1]; part[row * 4 + 2] = acc3[2]; }
        epi_bar_sync();
        if (half == 0) {
#pragma unroll
          for (int o = 0; o < 3; ++o) res[o] = acc3[o] + part[row * 4 + o];
        }
        epi_bar_sync();  // part[] is reused by the next dot epilogue
        ++lc;
      };

      const int n_it = k2 ? n_pass : (n_tiles > (int)blockIdx.x ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0);
      int tile = blockIdx.x;
      for (int it = 0; it < n_it; ++it, tile += gridDim.x) {  // (2-SM: a tile number past n_tiles is a phantom tile, s >= S)
        const long long s = (long long)tile * 128 + row;
        tile_now = tile;
        if (kMode == MODE_PROBE) {
          // one linear layer, fp32 dump (self-test of descriptors / swizzle / TMEM read-back)
          const uint32_t ab = lc & 1u;
          wait_acc(ab);
          const float inv = P.layers[0].inv_scale;
#pragma unroll 1
          for (int i = 0; i < 4; ++i) {
            const int c0 = (2 * i + half) * 32;
            uint32_t v[32];
            tmem_ld32_async(t_row + ab * 256u + c0, v);
            tmem_wait(v);
            if (s < S) {
#pragma unroll
              for (int j = 0; j < 32; ++j) P.feat_out[s * kHidden + c0 + j] = fmaf(__uint_as_float(v[j]), inv, P.bias[0][c0 + j]);
            }
          }
          release_acc(ab);
          ++lc;
          continue;
        }
        float* feat_row = (P.feat_out && s < S) ? P.feat_out + s * kHidden : nullptr;
        float sigma = 0.f;
        float res[3];
        // logical layer ll = l + layer_ofs (layer_ofs = 1: local_field.8 folded into the two layers that consume it, the tile's
        // input image then feeds shape_net.0' and channel_net.0' directly)
#pragma unroll 1
        for (int l = 0; l < n_layers; ++l) {
          const int ll = l + P.layer_ofs;
          if (ll == 1 || ll == 5) {
            // ll = 1: shape_net.0 + LeakyReLU, shape_net.2;  ll = 5: channel_net.6 + LeakyReLU, channel_net.8
            epilogue_dot(l, ll == 1 ? 1 : 3, res);
            if (ll == 1 && half == 0) {
              const float xs = res[0] + P.shape_out_b - 1.0f;
              sigma = xs > 20.f ? xs : log1pf(expf(xs));
            }
          } else {
            // ll = 0: local_field.8 (linear) -> feat;  ll = 2..4: channel_net.0,2,4 (layer 2 reads feat and overwrites it in place)
            epilogue_store(kWhole, l, ll == 0 ? 1.0f : 0.01f, ll == 0 ? feat_row : nullptr);
          }
        }
        if (half == 0 && s < S) {
          float rgb[3];
#pragma unroll
          for (int o = 0; o < 3; ++o) rgb[o] = 1.0f / (1.0f + expf(-(res[o] + P.chan_out_b[o])));
          P.rgbs[s] = make_float4(rgb[0], rgb[1], rgb[2], sigma);
        }
        if (et == 0) NPCD_TL(tl_it, 10);
        ++tl_it;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kCluster) cluster_sync_all();  // the partner may still multicast into / signal this CTA's shared memory until it is done too
  if (warp == 1) {
    __syncwarp();
    if (k2) tmem_dealloc_2sm(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// Pack an fp32 [256, k_in] nn.Linear weight into per-K-block pre-swizzled fp16 hi/lo tiles (the exact shared-memory image the
// MMA descriptors expect), multiplied by `scale` (a power of two).  perm[k'] = source column of packed column k' (or -1 = 0).
// format 0: fp16 hi tile + fp16 lo tile; format 1 (f16 + e4m3 x 2 scheme, tc_ptx.cuh): fp16 tile of v * 2^13 + one 8-bit tile whose
// row n is [Whi8 = e4m3(v * 2^5) of the 64 columns | Wlo8 = e4m3((v * 2^13 - W16) * 2^4) of the 64 columns]; format 2: the same values,
// 8-bit tile rows = four K = 32 steps of [Whi8 of 16 columns | Wlo8 of the same 16 columns] (A operand in tensor memory, kTS)
__device__ __forceinline__ void pack_weight_element(uint8_t* tile, int n, int kk, float v, int format, bool full_block) {
  const size_t off = (size_t)(n >> 3) * 1024 + (n & 7) * 128 + (((kk >> 3) ^ (n & 7)) << 4) + (kk & 7) * 2;
  if (format == 0) {
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    *reinterpret_cast<__half*>(tile + off) = hi;
    *reinterpret_cast<__half*>(tile + kTileBytesW + off) = lo;
  } else {
    const float vs = v * 8192.0f;
    const __half hi = __float2half_rn(vs);
    *reinterpret_cast<__half*>(tile + off) = hi;
    const uint32_t w8 = cvt_e4m3x2_f32(v * 32.0f, (vs - __half2float(hi)) * 16.0f);  // byte 0 = Whi8, byte 1 = Wlo8
    const int c8 = kk >> 3;
    uint8_t* t8 = tile + kTileBytesW;
    if (format == 2) {  // tensor-memory operand form: per 16 input features one K = 32 step, row bytes [Whi8 x 16 | Wlo8 x 16]
      const int t = kk >> 4;
      const size_t row = (size_t)(n >> 3) * 1024 + (n & 7) * 128;
      t8[row + (((2 * t) ^ (n & 7)) << 4) + (kk & 15)] = (uint8_t)(w8 & 0xffu);
      t8[row + (((2 * t + 1) ^ (n & 7)) << 4) + (kk & 15)] = (uint8_t)(w8 >> 8);
      return;
    }
    t8[swz8_lo(n, c8, full_block) + (kk & 7)] = (uint8_t)(w8 & 0xffu);
    t8[swz8_hi(n, c8, full_block) + (kk & 7)] = (uint8_t)(w8 >> 8);
  }
}

__global__ void k_pack_weights(const float* __restrict__ w, int k_in, const int* __restrict__ perm, int k_pad, float scale,
                               uint8_t* __restrict__ out, int format) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (n, k') pairs
  if (idx >= 256 * k_pad) return;
  const int n = idx / k_pad, kp = idx % k_pad;
  const int src = perm ? perm[kp] : (kp < k_in ? kp : -1);
  const float v = src >= 0 ? w[(size_t)n * k_in + src] * scale : 0.f;
  pack_weight_element(out + (size_t)(kp >> 6) * 2 * kTileBytesW, n, kp & 63, v, format, (kp | 63) < k_pad);
}

// Batched variant: every weight matrix of a training step (10 forward layers, 4 + 6 transposed dgrad operands) in ONE launch;
// a job may read its source transposed and keep only the first n_rows packed rows (layer 0 of the dgrad chain: 32 feature rows).
struct PackJobs {
  npcd_tc_pack_job j[24];
};
__global__ void __launch_bounds__(256) k_pack_weights_batched(const __grid_constant__ PackJobs J) {
  const npcd_tc_pack_job& job = J.j[blockIdx.y];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (n, k') pairs
  if (idx >= 256 * job.k_pad) return;
  const int n = idx / job.k_pad, kp = idx % job.k_pad;
  const int src = job.perm ? job.perm[kp] : (kp < job.k_in ? kp : -1);
  float v = 0.f;
  if (src >= 0 && n < job.n_rows) v = (job.transpose ? job.w[(size_t)src * job.ld + n] : job.w[(size_t)n * job.ld + src]) * job.scale;
  pack_weight_element((uint8_t*)job.out + (size_t)(kp >> 6) * 2 * kTileBytesW, n, kp & 63, v, job.format, (kp | 63) < job.k_pad);
}

// fp32 rows [n,256] <-> the pre-split operand image (per 128-row tile: 4 K-blocks x (hi 16 KB, lo 16 KB), SWIZZLE_128B)
template <bool kF8>
__global__ void k_rows_to_image(const float* __restrict__ x, long long n, uint8_t* __restrict__ img) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // (row, 8-column chunk)
  if (i >= n * 32) return;
  const long long s = i >> 5;
  const int c8 = (int)(i & 31);
  const float4 a = __ldg(reinterpret_cast<const float4*>(x + s * kHidden + c8 * 8));
  const float4 b = __ldg(reinterpret_cast<const float4*>(x + s * kHidden + c8 * 8 + 4));
  uint8_t* kbp = img + (size_t)(s >> 7) * kImgTileBytes + (size_t)(c8 >> 3) * (2 * kTileBytesA);
  if (!kF8) {
    const float y[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint4 hi, lo;
    split8(y, hi, lo);
    uint8_t* p = kbp + swz((int)(s & 127), c8 & 7);
    *reinterpret_cast<uint4*>(p) = hi;
    *reinterpret_cast<uint4*>(p + kTileBytesA) = lo;
  } else {
    const float k = kF8ActScale;
    const float y[8] = {a.x * k, a.y * k, a.z * k, a.w * k, b.x * k, b.y * k, b.z * k, b.w * k};
    uint4 hi;
    uint2 lo8, hi8;
    split8_f8(y, hi, lo8, hi8);
    *reinterpret_cast<uint4*>(kbp + swz((int)(s & 127), c8 & 7)) = hi;
    *reinterpret_cast<uint2*>(kbp + kTileBytesA + swz8_lo((int)(s & 127), c8 & 7)) = lo8;
    *reinterpret_cast<uint2*>(kbp + kTileBytesA + swz8_hi((int)(s & 127), c8 & 7)) = hi8;
  }
}

// (f8 scheme: the value the tensor cores see through the fp16 and lo8 parts, (A16 + lo8 / 2^8) / 2^3)
template <bool kF8>
__global__ void k_image_to_rows(const uint8_t* __restrict__ img, long long n, float* __restrict__ x) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 32) return;
  const long long s = i >> 5;
  const int c8 = (int)(i & 31);
  const uint8_t* kbp = img + (size_t)(s >> 7) * kImgTileBytes + (size_t)(c8 >> 3) * (2 * kTileBytesA);
  const uint8_t* p = kbp + swz((int)(s & 127), c8 & 7);
  const uint4 hi = *reinterpret_cast<const uint4*>(p);
  const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w};
  float y[8];
  if (!kF8) {
    const uint4 lo = *reinterpret_cast<const uint4*>(p + kTileBytesA);
    const uint32_t l[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 fh = __half22float2(*reinterpret_cast<const __half2*>(&h[j]));
      const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&l[j]));
      y[2 * j] = fh.x + fl.x;
      y[2 * j + 1] = fh.y + fl.y;
    }
  } else {
    const uint2 lo8 = *reinterpret_cast<const uint2*>(kbp + kTileBytesA + swz8_lo((int)(s & 127), c8 & 7));
    const uint32_t l[2] = {lo8.x, lo8.y};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 fh = __half22float2(*reinterpret_cast<const __half2*>(&h[j]));
      const uint16_t two = (uint16_t)(l[j >> 1] >> (16 * (j & 1)));
      uint32_t h2;
      asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(h2) : "h"(two));
      const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&h2));
      y[2 * j] = (fh.x + fl.x * (1.0f / 256.0f)) * (1.0f / kF8ActScale);
      y[2 * j + 1] = (fh.y + fl.y * (1.0f / 256.0f)) * (1.0f / kF8ActScale);
    }
  }
  *reinterpret_cast<float4*>(x + s * kHidden + c8 * 8) = make_float4(y[0], y[1], y[2], y[3]);
  *reinterpret_cast<float4*>(x + s * kHidden + c8 * 8 + 4) = make_float4(y[4], y[5], y[6], y[7]);
}

// ---- dense pair packing -------------------------------------------------------------------------------------------------
struct NbrCount {
  const int* nbr_idx;
  const long long* n_samples_dev;
  __device__ int operator()(long long s) const {
    if (s >= *n_samples_dev) return 0;
    const int4 a = __ldg(reinterpret_cast<const int4*>(nbr_idx + s * kK));
    const int4 b = __ldg(reinterpret_cast<const int4*>(nbr_idx + s * kK + 4));
    return (a.x >= 0) + (a.y >= 0) + (a.z >= 0) + (a.w >= 0) + (b.x >= 0) + (b.y >= 0) + (b.z >= 0) + (b.w >= 0);
  }
};

__global__ void k_zero_int(int* p) { p[0] = 0; }

// Dense pair tiles: a tile is a maximal run of WHOLE samples whose pairs fit the 128 rows of one MMA tile (greedy packing: ~125.4 used
// rows per tile on the benchmark clouds against 121 for fixed 121-offset windows, i.e. 3.5 % fewer tensor-core tiles).  The greedy
// chain is sequential, so it is cut every kTileBlock samples (a block always starts a new tile: one short tile per ~50): one thread
// walks each block (7-step binary search per tile over pair_off), first counting, then -- after a scan of the per-block counts --
// writing tile_start.  Every sample has 1..8 pairs, so every tile but the last of a block holds >= 121 rows.
constexpr int kTileBlock = 1024;

template <bool kWrite>
__global__ void k_tile_walk(const int* __restrict__ pair_off, const long long* __restrict__ n_samples_dev, long long capacity,
                            long long n_blocks, int* __restrict__ blk, int* __restrict__ tile_start) {
  const long long S = min(*n_samples_dev, capacity);
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_blocks) return;
  const long long s_begin = b * kTileBlock, s_end = min(S, s_begin + kTileBlock);
  const int base = kWrite ? blk[b] : 0;
  int n = 0;
  long long s = s_begin;
  while (s < s_end) {
    if (kWrite) tile_start[base + n] = (int)s;
    ++n;
    const int p0 = __ldg(pair_off + s);
    long long lo = s + 1, hi = min(s + 128, s_end);  // largest x with pair_off[x] - p0 <= 128 (x = s + 1 always qualifies)
    while (lo < hi) {
      const long long mid = (lo + hi + 1) >> 1;
      if (__ldg(pair_off + mid) - p0 <= 128) lo = mid; else hi = mid - 1;
    }
    s = lo;
  }
  if (!kWrite) blk[b] = n;
}

// exclusive scan of the per-block tile counts (single CTA), totals for the consumers
__global__ void __launch_bounds__(1024) k_tile_scan(int* __restrict__ blk, long long n_blocks, const long long* __restrict__ n_samples_dev,
                                                    long long capacity, int* __restrict__ tile_start, int* __restrict__ n_tiles_dev,
                                                    long long* __restrict__ rows_dev) {
  __shared__ int warp_sum[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long c0 = 0; c0 < n_blocks; c0 += blockDim.x) {
    const long long i = c0 + threadIdx.x;
    const int v = i < n_blocks ? blk[i] : 0;
    int incl = v;
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sum[lane];
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += u;
      }
      warp_sum[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const int carry = carry_s;
    if (i < n_blocks) blk[i] = carry + (warp ? warp_sum[warp - 1] : 0) + incl - v;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + warp_sum[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int total = carry_s;
    *n_tiles_dev = total;
    tile_start[total] = (int)min(*n_samples_dev, capacity);
    if (rows_dev) *rows_dev = (long long)total * 128;  // training stash: rows of the per-tile operand images
  }
}

}  // namespace tc
}  // namespace npcd

using namespace npcd;

extern "C" int npcd_tc_pack_weights(const float* w, int k_in, const int* perm, int k_pad, float scale, void* out, void* stream) {
  NPCD_CHECK_ARG(w && out, "null pointer");
  NPCD_CHECK_ARG(k_in > 0 && k_pad > 0 && k_pad % 16 == 0 && k_pad <= 256 && (perm || k_in <= k_pad), "bad sizes");
  const int n = 256 * k_pad;
  tc::k_pack_weights<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, k_in, perm, k_pad, scale, (uint8_t*)out, 0);
  return check_launch("npcd_tc_pack_weights");
}

extern "C" int npcd_tc_pack_weights_f8(const float* w, int k_in, const int* perm, int k_pad, float scale, void* out, void* stream) {
  NPCD_CHECK_ARG(w && out, "null pointer");
  NPCD_CHECK_ARG(k_in > 0 && k_pad > 0 && k_pad % 32 == 0 && k_pad <= 256 && (perm || k_in <= k_pad), "bad sizes");
  const int n = 256 * k_pad;
  tc::k_pack_weights<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, k_in, perm, k_pad, scale, (uint8_t*)out, 1);
  return check_launch("npcd_tc_pack_weights_f8");
}

extern "C" int npcd_tc_pack_weights_batched(const npcd_tc_pack_job* jobs, int n_jobs, void* stream) {
  NPCD_CHECK_ARG(jobs && n_jobs >= 0 && n_jobs <= 24, "at most 24 jobs per call");
  if (n_jobs == 0) return 0;
  tc::PackJobs J;
  memset(&J, 0, sizeof(J));
  for (int i = 0; i < n_jobs; ++i) {
    const npcd_tc_pack_job& j = jobs[i];
    NPCD_CHECK_ARG(j.w && j.out, "null pointer");
    NPCD_CHECK_ARG(j.k_in > 0 && j.k_pad > 0 && j.k_pad % 16 == 0 && j.k_pad <= 256 && (j.perm || j.k_in <= j.k_pad), "bad sizes");
    NPCD_CHECK_ARG(j.n_rows > 0 && j.n_rows <= 256 && j.ld > 0 && !(j.perm && j.transpose), "bad job");
    NPCD_CHECK_ARG(j.format == 0 || (j.format == 1 && j.k_pad % 32 == 0) || (j.format == 2 && j.k_pad % 64 == 0), "bad format");
    J.j[i] = j;
  }
  tc::k_pack_weights_batched<<<dim3(256, n_jobs), 256, 0, (cudaStream_t)stream>>>(J);
  return check_launch("npcd_tc_pack_weights_batched");
}

extern "C" int npcd_tc_rows_to_image(const float* rows, long long n, void* image, void* stream) {
  NPCD_CHECK_ARG(n >= 0 && (n == 0 || (rows && image)), "bad arguments");
  if (n == 0) return 0;
  tc::k_rows_to_image<false><<<(unsigned)((n * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rows, n, (uint8_t*)image);
  return check_launch("npcd_tc_rows_to_image");
}

extern "C" int npcd_tc_rows_to_image_f8(const float* rows, long long n, void* image, void* stream) {
  NPCD_CHECK_ARG(n >= 0 && (n == 0 || (rows && image)), "bad arguments");
  if (n == 0) return 0;
  tc::k_rows_to_image<true><<<(unsigned)((n * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rows, n, (uint8_t*)image);
  return check_launch("npcd_tc_rows_to_image_f8");
}

extern "C" int npcd_tc_image_to_rows(const void* image, long long n, float* rows, void* stream) {
  NPCD_CHECK_ARG(n >= 0 && (n == 0 || (rows && image)), "bad arguments");
  if (n == 0) return 0;
  tc::k_image_to_rows<false><<<(unsigned)((n * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)image, n, rows);
  return check_launch("npcd_tc_image_to_rows");
}

extern "C" int npcd_tc_image_to_rows_f8(const void* image, long long n, float* rows, void* stream) {
  NPCD_CHECK_ARG(n >= 0 && (n == 0 || (rows && image)), "bad arguments");
  if (n == 0) return 0;
  tc::k_image_to_rows<true><<<(unsigned)((n * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)image, n, rows);
  return check_launch("npcd_tc_image_to_rows_f8");
}

// workspace layout: [operand image | pair_off (capacity+1) | tile_start (max_tiles+2) | n_tiles (1, padded) | cub scratch]
namespace {
struct TcWorkspace {
  size_t img_off, img_bytes, pair_off, tile_off, ntiles_off, blk_off, cub_off, cub_bytes, total;
  long long max_tiles;
};
inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
// every greedy tile holds >= kPackRows pair rows except the last tile of each kTileBlock-sample block
inline long long max_pair_tiles(long long capacity) { return capacity * kK / tc::kPackRows + capacity / tc::kTileBlock + 4; }

int tc_workspace_layout(long long capacity, TcWorkspace* w) {
  w->img_off = 0;
  w->img_bytes = (size_t)((capacity + 127) / 128) * tc::kImgTileBytes;
  w->pair_off = align256(w->img_off + w->img_bytes);
  w->max_tiles = max_pair_tiles(capacity);
  w->tile_off = align256(w->pair_off + (size_t)(capacity + 1) * sizeof(int));
  w->ntiles_off = align256(w->tile_off + (size_t)(w->max_tiles + 2) * sizeof(int));
  w->blk_off = w->ntiles_off + 256;
  w->cub_off = align256(w->blk_off + (size_t)(capacity / tc::kTileBlock + 2) * sizeof(int));
  size_t tmp = 0;
  tc::NbrCount op{nullptr, nullptr};
  cub::CountingInputIterator<long long> cnt(0);
  cub::TransformInputIterator<int, tc::NbrCount, cub::CountingInputIterator<long long>> it(cnt, op);
  cudaError_t e = cub::DeviceScan::InclusiveSum(nullptr, tmp, it, (int*)nullptr, capacity > 0 ? capacity : 1);
  if (e != cudaSuccess) {
    set_error("npcd_field_tc_workspace_bytes: %s", cudaGetErrorString(e));
    return 2;
  }
  w->cub_bytes = tmp + 256;
  w->total = w->cub_off + w->cub_bytes;
  return 0;
}

// f8 scheme (tc_ptx.cuh): the accumulator holds 2^16 x the pre-scaled product, and a layer whose output becomes the next operand
// (ACT / LINEAR / AGG epilogues) produces it times kF8ActScale -- both folded into the epilogue's scale and bias here.
void fill_layer(tc::Params& P, int i, const npcd_tc_layer& src, int epi, bool f8 = false) {
  P.layers[i].w = (const uint8_t*)src.packed_w;
  P.layers[i].inv_scale = src.inv_scale;
  P.layers[i].ksteps = src.k_pad / 16;
  P.layers[i].epi = epi;
  memcpy(P.bias[i], src.bias, sizeof(float) * 256);
  if (f8) {
    const float out_mult = (epi == tc::EPI_ACT || epi == tc::EPI_LINEAR || epi == tc::EPI_AGG) ? tc::kF8ActScale : 1.0f;
    P.layers[i].inv_scale = src.inv_scale * tc::kF8AccScaleInv * out_mult;
    for (int c = 0; c < 256; ++c) P.bias[i][c] *= out_mult;
  }
}

long long* g_timeline = nullptr;
long long* g_timeline_heads = nullptr;

template <int kMode, bool kF8 = false, bool kTS = false>
int launch_tc(const tc::Params& P_in, long long tiles, int num_sms, cudaStream_t st, const char* what) {
  static thread_local tc::Params P;
  P = P_in;
  P.timeline = kMode == tc::MODE_PAIR ? g_timeline : (kMode == tc::MODE_HEADS ? g_timeline_heads : nullptr);
  cudaError_t e = cudaFuncSetAttribute(tc::k_field_tc<kMode, kF8, kTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemTotal);
  if (e != cudaSuccess) {
    set_error("%s: cannot opt in to %d bytes of shared memory: %s", what, tc::kSmemTotal, cudaGetErrorString(e));
    return 2;
  }
  if (num_sms <= 0) num_sms = 148;
  unsigned grid = (unsigned)(tiles < num_sms ? (tiles > 0 ? tiles : 1) : num_sms);
  if (tc::kCluster) grid = (grid + 1u) & ~1u;  // whole 2-CTA clusters
  tc::k_field_tc<kMode, kF8, kTS><<<grid, tc::threads_for<kMode>(), tc::kSmemTotal, st>>>(P);
  return check_launch(what);
}
}  // namespace

namespace {
// dense packing (pair_off = exclusive scan of the neighbour counts, greedy tile starts) + the pair kernel
template <int kMode, bool kF8 = false, bool kTS = false>
int pair_stage(const int* nbr_idx, const float* sample_pos, const float* kp_pos, const float* kp_feat, const long long* n_samples_dev,
               long long capacity, const npcd_mlp_tc_weights* W, const TcWorkspace& ws, uint8_t* base,
               const npcd_pair_stash_layout* layout, uint8_t* stash, int* error_flag, int num_sms, cudaStream_t st,
               int no_wcorr = 0) {
  uint8_t* img = base + ws.img_off;
  int* pair_off = (int*)(base + ws.pair_off);
  int* tile_start = (int*)(base + ws.tile_off);
  int* n_tiles_dev = (int*)(base + ws.ntiles_off);
  tc::k_zero_int<<<1, 1, 0, st>>>(pair_off);
  tc::NbrCount op{nbr_idx, n_samples_dev};
  cub::CountingInputIterator<long long> cnt(0);
  cub::TransformInputIterator<int, tc::NbrCount, cub::CountingInputIterator<long long>> it(cnt, op);
  size_t tmp = ws.cub_bytes;
  cudaError_t e = cub::DeviceScan::InclusiveSum(base + ws.cub_off, tmp, it, pair_off + 1, capacity, st);
  if (e != cudaSuccess) {
    set_error("npcd_field_tc_fwd: scan: %s", cudaGetErrorString(e));
    return 2;
  }
  int* blk = (int*)(base + ws.blk_off);
  const long long n_blocks = capacity / tc::kTileBlock + 1;
  tc::k_tile_walk<false><<<(unsigned)((n_blocks + 63) / 64), 64, 0, st>>>(pair_off, n_samples_dev, capacity, n_blocks, blk, tile_start);
  tc::k_tile_scan<<<1, 1024, 0, st>>>(blk, n_blocks, n_samples_dev, capacity, tile_start, n_tiles_dev,
                                      stash ? (long long*)(stash + layout->rows_dev) : nullptr);
  tc::k_tile_walk<true><<<(unsigned)((n_blocks + 63) / 64), 64, 0, st>>>(pair_off, n_samples_dev, capacity, n_blocks, blk, tile_start);
  int rc = check_launch("npcd_field_tc_fwd(pack)");
  if (rc) return rc;
  static thread_local tc::Params P;  // ~11 KB: keep it off the stack
  memset(&P, 0, sizeof(P));
  for (int i = 0; i < 4; ++i) fill_layer(P, i, W->pair[i], i < 3 ? tc::EPI_ACT : tc::EPI_AGG, kF8);
  P.n_layers = 4;
  P.nbr_idx = nbr_idx; P.sample_pos = (const float4*)sample_pos; P.kp_pos = kp_pos; P.kp_feat = kp_feat;
  P.pair_off = pair_off; P.tile_start = tile_start; P.n_tiles_dev = n_tiles_dev; P.img = img;
  P.n_samples_dev = n_samples_dev; P.capacity = capacity; P.error_flag = error_flag;
  P.no_wcorr = no_wcorr;
  if (stash) {
    for (int l = 0; l < 4; ++l) {
      P.stash_x[l] = stash + layout->x[l];
      P.stash_mask[l] = (uint32_t*)(stash + layout->mask[l]);
    }
    P.stash_wn = (float*)(stash + layout->wn);
    P.stash_idx = (int*)(stash + layout->idx);
    P.stash_samp = (int*)(stash + layout->samp);
  }
  return launch_tc<kMode, kF8, kTS>(P, ws.max_tiles, num_sms, st, "npcd_field_tc_fwd(pair)");
}
}  // namespace

namespace {
template <int kMode, bool kF8 = false, bool kTS = false>
int heads_stage(const npcd_mlp_tc_weights* W, uint8_t* img, float* rgbs, float* feat_out, const long long* n_samples_dev,
                long long capacity, const npcd_pair_stash_layout* layout, uint8_t* stash, int* error_flag, int num_sms, cudaStream_t st,
                bool folded = false, int no_wcorr = 0) {
  static thread_local tc::Params P;
  memset(&P, 0, sizeof(P));
  P.no_wcorr = no_wcorr;
  const int o = folded ? 1 : 0;  // folded: W->shape / W->chan[0] already contain local_field.8 (W' = W W_8, b' = W b_8 + b)
  if (!folded) fill_layer(P, 0, W->agg, tc::EPI_LINEAR, kF8);
  fill_layer(P, 1 - o, W->shape, tc::EPI_DOT1, kF8);
  for (int i = 0; i < 3; ++i) fill_layer(P, 2 + i - o, W->chan[i], tc::EPI_ACT, kF8);
  fill_layer(P, 5 - o, W->chan[3], tc::EPI_DOT3, kF8);
  P.n_layers = 6 - o;
  P.layer_ofs = o;
  P.img = img; P.rgbs = (float4*)rgbs; P.feat_out = feat_out;
  memcpy(P.shape_out_w, W->shape_out_w, sizeof(float) * 256);
  memcpy(P.chan_out_w, W->chan_out_w, sizeof(float) * 3 * 256);
  P.shape_out_b = W->shape_out_b[0];
  memcpy(P.chan_out_b, W->chan_out_b, sizeof(float) * 3);
  P.n_samples_dev = n_samples_dev; P.capacity = capacity; P.error_flag = error_flag;
  if (stash) {
    for (int i = 0; i < 6; ++i) P.hstash_x[i] = stash + layout->hx[i];
    for (int i = 0; i < 5; ++i) P.hstash_mask[i] = (uint32_t*)(stash + layout->hmask[i]);
  }
  return launch_tc<kMode, kF8, kTS>(P, (capacity + 127) / 128, num_sms, st, "npcd_field_tc_fwd(heads)");
}
}  // namespace

// development aid: the inference pair kernel's CTA 0 records clock64() at its phase boundaries into buf [64 tiles][32] (null: off)
extern "C" int npcd_debug_set_timeline(void* buf) {
  g_timeline = (long long*)buf;
  return 0;
}
extern "C" int npcd_debug_set_timeline_heads(void* buf) {
  g_timeline_heads = (long long*)buf;
  return 0;
}

extern "C" int npcd_field_tc_workspace_bytes(long long capacity, size_t* bytes) {
  NPCD_CHECK_ARG(bytes && capacity >= 0 && capacity < (1ll << 27), "bad arguments (capacity must be < 2^27 samples per launch)");
  TcWorkspace w;
  int rc = tc_workspace_layout(capacity, &w);
  if (rc) return rc;
  *bytes = w.total;
  return 0;
}

extern "C" int npcd_field_tc_fwd(const int* nbr_idx, const float* sample_pos, const float* kp_pos, const float* kp_feat,
                                 const long long* n_samples_dev, long long capacity, const npcd_mlp_tc_weights* W,
                                 void* workspace, size_t workspace_bytes, float* rgbs, float* feat_out, int stages,
                                 int* error_flag, int num_sms, void* stream) {
  NPCD_CHECK_ARG(n_samples_dev && W, "null pointer");
  NPCD_CHECK_ARG(capacity >= 0 && capacity < (1ll << 27), "bad capacity (must be < 2^27 samples per launch)");
  if (capacity == 0) return 0;
  NPCD_CHECK_ARG(nbr_idx && sample_pos && kp_pos && kp_feat && workspace && rgbs, "null pointer");
  NPCD_CHECK_ARG(W->feat_dim == 32, "the tensor-core field kernel is specialised for feat_dim = 32 (configs/npcd_srncars.yaml:6)");
  TcWorkspace ws;
  int rc = tc_workspace_layout(capacity, &ws);
  if (rc) return rc;
  NPCD_CHECK_ARG(workspace_bytes >= ws.total, "workspace too small (npcd_field_tc_workspace_bytes)");
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* base = (uint8_t*)workspace;
  uint8_t* img = base + ws.img_off;
  int* pair_off = (int*)(base + ws.pair_off);
  int* tile_start = (int*)(base + ws.tile_off);
  int* n_tiles_dev = (int*)(base + ws.ntiles_off);
  const bool f8 = (stages & 8) != 0;  // W was packed with format 1 (f16 + e4m3 x 2 operand scheme)
  const int no_wcorr = (stages & 16) ? 1 : 0;  // ... and only the activation-rounding correction is issued ("f16+e4m3")
  NPCD_CHECK_ARG(!no_wcorr || f8, "stages bit 4 (single correction product) needs bit 3 (f16 + e4m3 operands)");
  const bool ts = (stages & 32) != 0;  // W->pair[1..3] were packed with format 2: A operand of those layers in tensor memory
  NPCD_CHECK_ARG(!ts || (f8 && !no_wcorr), "stages bit 5 (tensor-memory operand form) needs bit 3 and excludes bit 4");
  if (stages & 1) {
    rc = ts ? pair_stage<tc::MODE_PAIR, true, true>(nbr_idx, sample_pos, kp_pos, kp_feat, n_samples_dev, capacity, W, ws, base,
                                                                 nullptr, nullptr, error_flag, num_sms, st, 0)
       : f8 ? pair_stage<tc::MODE_PAIR, true>(nbr_idx, sample_pos, kp_pos, kp_feat, n_samples_dev, capacity, W, ws, base, nullptr, nullptr,
                                              error_flag, num_sms, st, no_wcorr)
            : pair_stage<tc::MODE_PAIR>(nbr_idx, sample_pos, kp_pos, kp_feat, n_samples_dev, capacity, W, ws, base, nullptr, nullptr,
                                        error_flag, num_sms, st);
    if (rc) return rc;
  }
  NPCD_CHECK_ARG(!(stages & 4) || !feat_out, "the folded heads stage has no local_field.8 output to return");
  NPCD_CHECK_ARG(!ts || !(stages & 2), "stages bit 5: the heads stage has a tensor-memory form only with local_field.8 folded (bit 2)");
  if (ts && (stages & 4))  // W->chan[1..3] in format 2
    rc = heads_stage<tc::MODE_HEADS, true, true>(W, img, rgbs, feat_out, n_samples_dev, capacity, nullptr, nullptr, error_flag,
                                                             num_sms, st, true, 0);
  else if (stages & 6)
    rc = f8 ? heads_stage<tc::MODE_HEADS, true>(W, img, rgbs, feat_out, n_samples_dev, capacity, nullptr, nullptr, error_flag, num_sms, st,
                                                (stages & 4) != 0, no_wcorr)
            : heads_stage<tc::MODE_HEADS>(W, img, rgbs, feat_out, n_samples_dev, capacity, nullptr, nullptr, error_flag, num_sms, st,
                                          (stages & 4) != 0);
  return rc;
}

// Probe / self-test: out[s, :] = x[s, :] @ W^T + b for one packed 256x256 layer (validates descriptors, swizzle, TMEM readback).
// `image` is the pre-split operand image of x (npcd_tc_rows_to_image), ceil(capacity / 128) * 128 KB.
extern "C" int npcd_tc_linear_probe(const void* image, const long long* n_rows_dev, long long capacity, const npcd_tc_layer* layer,
                                    float* out, int* error_flag, int num_sms, void* stream) {
  NPCD_CHECK_ARG(image && n_rows_dev && layer && out, "null pointer");
  NPCD_CHECK_ARG(capacity > 0, "bad capacity");
  static thread_local tc::Params P;
  memset(&P, 0, sizeof(P));
  fill_layer(P, 0, *layer, tc::EPI_DUMP);
  P.n_layers = 1;
  P.img = (uint8_t*)const_cast<void*>(image);
  P.feat_out = out;
  P.n_samples_dev = n_rows_dev; P.capacity = capacity; P.error_flag = error_flag;
  return launch_tc<tc::MODE_PROBE>(P, (capacity + 127) / 128, num_sms, (cudaStream_t)stream, "npcd_tc_linear_probe");
}

// same for the f16 + e4m3 x 2 operand scheme: `image` = npcd_tc_rows_to_image_f8(x), layer packed by npcd_tc_pack_weights_f8
extern "C" int npcd_tc_linear_probe_f8(const void* image, const long long* n_rows_dev, long long capacity, const npcd_tc_layer* layer,
                                       float* out, int* error_flag, int num_sms, void* stream) {
  NPCD_CHECK_ARG(image && n_rows_dev && layer && out, "null pointer");
  NPCD_CHECK_ARG(capacity > 0, "bad capacity");
  static thread_local tc::Params P;
  memset(&P, 0, sizeof(P));
  fill_layer(P, 0, *layer, tc::EPI_DUMP, true);
  P.n_layers = 1;
  P.img = (uint8_t*)const_cast<void*>(image);
  P.feat_out = out;
  P.n_samples_dev = n_rows_dev; P.capacity = capacity; P.error_flag = error_flag;
  return launch_tc<tc::MODE_PROBE, true>(P, (capacity + 127) / 128, num_sms, (cudaStream_t)stream, "npcd_tc_linear_probe_f8");
}

// ---- training forward of the pair stage: same kernel, plus the stash the fused backward needs ------------------------------
extern "C" int npcd_pair_stash_layout_for(long long capacity, npcd_pair_stash_layout* out) {
  NPCD_CHECK_ARG(out && capacity >= 0 && capacity < (1ll << 27), "bad arguments");
  const long long max_tiles = max_pair_tiles(capacity);  // as in tc_workspace_layout (host arithmetic only: no device needed)
  const size_t tiles = (size_t)max_tiles;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align256(off + bytes); return o; };
  out->max_tiles = max_tiles;
  out->x[0] = take(tiles * 2 * (2 * tc::kTileBytesA));
  for (int l = 1; l < 4; ++l) out->x[l] = take(tiles * 4 * (2 * tc::kTileBytesA));
  for (int l = 0; l < 4; ++l) out->dp[l] = take(tiles * 4 * (2 * tc::kTileBytesA));
  for (int l = 0; l < 4; ++l) out->mask[l] = take(tiles * 128 * 8 * sizeof(uint32_t));
  out->wn = take(tiles * 128 * sizeof(float));
  out->idx = take(tiles * 128 * sizeof(int));
  out->samp = take(tiles * 128 * sizeof(int));
  out->rows_dev = take(256);
  const size_t ht = (size_t)((capacity + 127) / 128);
  out->h_tiles = (long long)ht;
  for (int i = 0; i < 6; ++i) out->hx[i] = take(ht * 4 * (2 * tc::kTileBytesA));
  for (int i = 0; i < 6; ++i) out->hdp[i] = take(ht * 4 * (2 * tc::kTileBytesA));
  for (int i = 0; i < 5; ++i) out->hmask[i] = take(ht * 128 * 8 * sizeof(uint32_t));
  out->g4 = take(ht * 128 * 4 * sizeof(float));
  out->d_agg = take(ht * 128 * (size_t)kHidden * sizeof(float));
  out->total = off;
  return 0;
}

extern "C" int npcd_pair_tc_train_fwd(const int* nbr_idx, const float* sample_pos, const float* kp_pos, const float* kp_feat,
                                      const long long* n_samples_dev, long long capacity, const npcd_mlp_tc_weights* W,
                                      void* workspace, size_t workspace_bytes, const npcd_pair_stash_layout* layout, void* stash,
                                      size_t stash_bytes, int* error_flag, int num_sms, void* stream) {
  NPCD_CHECK_ARG(n_samples_dev && W && layout && stash, "null pointer");
  NPCD_CHECK_ARG(capacity > 0 && capacity < (1ll << 27), "bad capacity (0 < capacity < 2^27 samples per launch)");
  NPCD_CHECK_ARG(nbr_idx && sample_pos && kp_pos && kp_feat && workspace, "null pointer");
  NPCD_CHECK_ARG(W->feat_dim == 32, "the tensor-core field kernel is specialised for feat_dim = 32 (configs/npcd_srncars.yaml:6)");
  TcWorkspace ws;
  int rc = tc_workspace_layout(capacity, &ws);
  if (rc) return rc;
  NPCD_CHECK_ARG(workspace_bytes >= ws.total, "workspace too small (npcd_field_tc_workspace_bytes)");
  NPCD_CHECK_ARG(layout->max_tiles == ws.max_tiles && stash_bytes >= layout->total, "stash layout does not match the capacity");
  return pair_stage<tc::MODE_PAIR_TRAIN>(nbr_idx, sample_pos, kp_pos, kp_feat, n_samples_dev, capacity, W, ws, (uint8_t*)workspace,
                                         layout, (uint8_t*)stash, error_flag, num_sms, (cudaStream_t)stream);
}

extern "C" int npcd_field_tc_train_fwd(const int* nbr_idx, const float* sample_pos, const float* kp_pos, const float* kp_feat,
                                       const long long* n_samples_dev, long long capacity, const npcd_mlp_tc_weights* W,
                                       void* workspace, size_t workspace_bytes, const npcd_pair_stash_layout* layout, void* stash,
                                       size_t stash_bytes, float* rgbs, int* error_flag, int num_sms, void* stream) {
  NPCD_CHECK_ARG(rgbs && workspace && layout, "null pointer");
  NPCD_CHECK_ARG(capacity > 0 && (capacity + 127) / 128 == layout->h_tiles, "stash layout does not match the capacity");
  cudaStream_t st = (cudaStream_t)stream;
  // rows of the last 128-sample tile beyond `capacity` are never written by the pair stage: zero them so that the stashed layer
  // inputs of those rows are finite (they meet zero gradients in the weight-gradient GEMMs)
  cudaError_t e = cudaMemsetAsync((uint8_t*)workspace + (size_t)(layout->h_tiles - 1) * tc::kImgTileBytes, 0, tc::kImgTileBytes, st);
  if (e != cudaSuccess) {
    set_error("npcd_field_tc_train_fwd: %s", cudaGetErrorString(e));
    return 2;
  }
  int rc = npcd_pair_tc_train_fwd(nbr_idx, sample_pos, kp_pos, kp_feat, n_samples_dev, capacity, W, workspace, workspace_bytes, layout,
                                  stash, stash_bytes, error_flag, num_sms, stream);
  if (rc) return rc;
  return heads_stage<tc::MODE_HEADS_TRAIN>(W, (uint8_t*)workspace, rgbs, nullptr, n_samples_dev, capacity, layout, (uint8_t*)stash,
                                           error_flag, num_sms, st);
}
