// Internal interface of the tensor-core first stage of the scan (tcscan.cu), used by scan.cu.
#pragma once
#include "common.cuh"

namespace mepp {

constexpr int kTcMaxTiles = 16;           // N-tiles (of 8, 16 or 32 scan units = 64, 128 or 256 columns) per launch
constexpr int kTcUnitsPerTile = 32;
constexpr int kTcRowsPerTile = 128;       // windows (at a stride of four bases) per M-tile
constexpr int kTcBasesPerTile = 4 * kTcRowsPerTile;
constexpr int kTcEpiWarps = 16;           // epilogue warps = candidate lists per CTA
constexpr int kTcMaxBBytes = 176 * 1024;  // weight operand resident in shared memory

// One N-tile of a launch: `n_units` scan units starting at unit `unit0` of the launch's unit list, 8 columns each
// (4 phases x 2 filters), K = 32 * ksteps one-hot channels (4 per base), weights at byte `b_off` of the operand.
struct TcTile {
  uint32_t b_off;
  uint16_t ksteps, bn, unit0, n_units;
};

struct TcLaunch {
  TcTile tiles[kTcMaxTiles];
  int n_tiles;
  int ks_max;          // largest ksteps of the launch (bytes of an A tile: 2048 + 32 * ks_max)
  uint32_t b_bytes;    // bytes of the weight operand of this launch
};

struct TcScanArgs {
  const uint8_t* onehot;      // flat one-hot stream, 4 bytes per base (A, C, G, T channels; N = 1,1,1,1; others x 4)
  int64_t n_mtiles;           // ceil(N * L / 512)
  const uint8_t* B;           // int8 weights of this launch in the UMMA K-major no-swizzle tiling
  uint2* cand;                // candidate lists [gridDim.x * kTcEpiWarps][list_cap]: {row, unit << 2 | phase}
  uint32_t* cand_count;       // [gridDim.x * kTcEpiWarps] (may exceed list_cap: overflow)
  uint32_t list_cap;
  int32_t* dbg_acc;           // optional: raw accumulators of M-tile 0 [128][dbg_ld]
  int dbg_ld;
  unsigned long long* dbg_clk;   // optional: cycle stamps of CTA 0's pipeline, [5][256] (bring-up / tuning only)
  int mode;                   // timing experiments only (MEPP_TC_MODE): 1 no epilogue work, 2 no MMAs, 4 loads but no tests
};

int tc_scan_grid();   // CTAs of a launch (one per SM)
int tc_units_per_tile();   // 16 (four accumulator stages of 128 columns, default) or 32 (two of 256)

// Fills the int8 weight operand of one launch from the float32 weights and biases of its units (device side: the
// thresholds may have been computed on the GPU).
int tc_build_weights(const TcLaunch& ln, const float* lut_dev, const float* bias_dev, const int32_t* mlen_dev,
                     const int32_t* nfilt_dev, const int32_t* units_dev, int unit_base, int n_units_total, uint8_t* B_dev,
                     cudaStream_t st);
int tc_scan_launch(const TcLaunch& ln, const TcScanArgs& a, cudaStream_t st);

}  // namespace mepp
