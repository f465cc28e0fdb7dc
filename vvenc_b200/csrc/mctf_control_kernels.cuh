// mctf_control_kernels.cuh -- the control of the MCTF motion search on the device: MCTF::motionEstimationLuma / estimateLumaLn (CommonLib/MCTF.cpp:1166-1397).
// The error tables come from mctf_error_packed_kernel / mctf_grid_kernel (mctf_affine_kernels.cuh); the kernels here build the candidate lists, replay the
// `error < best.error` chains in the reference's loop order and resolve the dependency on the upper and left neighbour (the prevLineX scheme of :1176, 1357-1386)
// with one warp per block row that waits for the row above, so that a whole pyramid level runs without the host looking at a number.
//   stage A  mctf_pred_cands_kernel + mctf_select_list_kernel : the 3x3 neighbourhood of the coarser level's field and the zero vector       (:1191-1214)
//   stage B  mctf_centre_kernel + grid + mctf_select_grid_kernel: integer grid around trunc(best / 16)                                      (:1216-1228)
//   stage C  the same pair, three times, around the running best with the centre skipped (doubleRes)                                        (:1229-1287)
//   stage D  mctf_wave_kernel: final vectors of the block above and of the block to the left                                                (:1288-1306)
//   stage E  mctf_final_kernel: error scaling with the block variance, rmsme (doubleRes)                                                    (:1308-1321)
// vvenc_b200/mctf_host.py holds the same replay on the host (the round-1 path, kept as the test oracle's driver).
#pragma once
#include "common.cuh"
#include "mctf_affine_kernels.cuh"

namespace vvb {

struct MctfGeom
{
  int width, height, bs, bxn, byn, n;      // level picture, block size, blocks per row / column (`blockX + 8 <= origWidth`, :1174, 1388)
  int prevW, prevH, factor;                // coarser level's field (0 x 0: none) and the vector scale between the levels
  int outW, outH;                          // field array the level writes into (entries no block writes keep the default vector 0,0)
};
struct MctfBest { int x, y, e; };

__device__ __forceinline__ void mctf_block_of( const MctfGeom& g, int i, int& X, int& Y, int& W, int& H )
{
  const int by = i / g.bxn, bx = i - by * g.bxn;
  X = bx * g.bs; Y = by * g.bs;
  W = min( g.bs, g.width - X ) & ~7; H = min( g.bs, g.height - Y ) & ~7;
}

__global__ void mctf_init_kernel( MctfGeom g, MctfBest* __restrict__ best, int* __restrict__ progress )
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if( i < g.n ) { best[i].x = 0; best[i].y = 0; best[i].e = 0x7fffffff; }          // MotionVector(): error = INT_LEAST32_MAX (MCTF.h:79)
  if( i <= g.byn ) progress[i] = 0;                                                 // [0] = row ticket, [1 + row] = blocks finished in that row
}

// candidate k of block i: k = 0..8 the coarser level's vector at (Y / 2bs + dy, X / 2bs + dx), dy outer, dx inner; k = 9 the zero vector.  Positions outside
// the coarser field are evaluated as the zero vector and ignored by the selection.
__global__ void mctf_pred_cands_kernel( MctfGeom g, const vvb_mctf_mv* __restrict__ prev, vvb_mctf_cand* __restrict__ cands )
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if( t >= g.n * 10 ) return;
  const int i = t / 10, k = t - i * 10;
  int X, Y, W, H; mctf_block_of( g, i, X, Y, W, H );
  vvb_mctf_cand c; c.x = X; c.y = Y; c.w = (uint16_t) W; c.h = (uint16_t) H; c.mvx = 0; c.mvy = 0;
  if( k < 9 )
  {
    const int ty = Y / ( 2 * g.bs ) + k / 3 - 1, tx = X / ( 2 * g.bs ) + k % 3 - 1;
    if( ty >= 0 && ty < g.prevH && tx >= 0 && tx < g.prevW ) { const vvb_mctf_mv p = prev[ty * g.prevW + tx]; c.mvx = p.x * g.factor; c.mvy = p.y * g.factor; }
  }
  cands[t] = c;
}

__global__ void mctf_select_list_kernel( MctfGeom g, const vvb_mctf_cand* __restrict__ cands, const int32_t* __restrict__ err, MctfBest* __restrict__ best )
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if( i >= g.n ) return;
  int X, Y, W, H; mctf_block_of( g, i, X, Y, W, H );
  MctfBest b = best[i];
  for( int k = 0; k < 10; k++ )
  {
    if( k < 9 )
    {
      const int ty = Y / ( 2 * g.bs ) + k / 3 - 1, tx = X / ( 2 * g.bs ) + k % 3 - 1;
      if( !( ty >= 0 && ty < g.prevH && tx >= 0 && tx < g.prevW ) ) continue;
    }
    const int e = err[i * 10 + k];
    if( e < b.e ) { b.e = e; b.x = cands[i * 10 + k].mvx; b.y = cands[i * 10 + k].mvy; }
  }
  best[i] = b;
}

// grid centre of every block: truncInt != 0 -> trunc( best / 16 ) * 16 (C division, `prevBest.x / m_motionVectorFactor`), else the running best; the block list
// handed to the grid kernel carries centre + shift (the lattice of the grid call may be offset against the offsets the selection visits)
__global__ void mctf_centre_kernel( MctfGeom g, const MctfBest* __restrict__ best, int truncInt, int shift, int2* __restrict__ centre, vvb_mctf_cand* __restrict__ blocks )
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if( i >= g.n ) return;
  int X, Y, W, H; mctf_block_of( g, i, X, Y, W, H );
  int cx = best[i].x, cy = best[i].y;
  if( truncInt ) { cx = ( cx / 16 ) * 16; cy = ( cy / 16 ) * 16; }
  centre[i] = make_int2( cx, cy );
  vvb_mctf_cand c; c.x = X; c.y = Y; c.w = (uint16_t) W; c.h = (uint16_t) H; c.mvx = cx + shift; c.mvy = cy + shift;
  blocks[i] = c;
}

// offsets o = off0 + k * delta (k = 0..count-1) in both directions, y outer, x inner, strictly smaller wins; the table holds the lattice (o - off0) / step
__global__ void mctf_select_grid_kernel( MctfGeom g, const int32_t* __restrict__ tab, int K1, int off0, int delta, int count, int step, int skipZero,
                                         const int2* __restrict__ centre, MctfBest* __restrict__ best )
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if( i >= g.n ) return;
  MctfBest b = best[i];
  const int2 c = centre[i];
  const int32_t* t = tab + (size_t) i * K1 * K1;
  const int stride = delta / step;
  for( int j = 0; j < count; j++ )
    for( int k = 0; k < count; k++ )
    {
      const int oy = off0 + j * delta, ox = off0 + k * delta;
      if( skipZero && ox == 0 && oy == 0 ) continue;
      const int e = t[j * stride * K1 + k * stride];
      if( e < b.e ) { b.e = e; b.x = c.x + ox; b.y = c.y + oy; }
    }
  best[i] = b;
}

// stage D: one warp per block row (rows are claimed through a ticket, so a waiting row always has the row above already running or finished)
__global__ void __launch_bounds__( 32 ) mctf_wave_kernel( const __grid_constant__ Plane orgPlane, const __grid_constant__ Plane refPlane, MctfGeom g, int tap4, int maxDim,
                                                          MctfBest* best, int* progress )
{
  extern __shared__ __align__( 16 ) uint32_t sWave[];
  const MctfSmem L = mctf_smem( maxDim );
  uint32_t* region = sWave; uint32_t* t2 = region + L.regionWords;
  const int lane = threadIdx.x;
  int by = 0;
  if( lane == 0 ) by = atomicAdd( &progress[0], 1 );
  by = __shfl_sync( 0xffffffffu, by, 0 );
  if( by >= g.byn ) return;
  volatile int* above = progress + by;                  // progress[1 + (by - 1)]
  MctfBest left = { 0, 0, 0 };
  for( int bx = 0; bx < g.bxn; bx++ )
  {
    const int i = by * g.bxn + bx;
    int X, Y, W, H; mctf_block_of( g, i, X, Y, W, H );
    MctfBest b;
    b.x = best[i].x; b.y = best[i].y; b.e = best[i].e;
    vvb_mctf_cand c; c.x = X; c.y = Y; c.w = (uint16_t) W; c.h = (uint16_t) H;
    if( by > 0 )
    {
      if( lane == 0 ) while( *above <= bx ) __nanosleep( 64 );
      __syncwarp();
      __threadfence();
      const MctfBest* up = best + i - g.bxn;
      c.mvx = __ldcg( &up->x ); c.mvy = __ldcg( &up->y );
      const int e = mctf_warp_error( orgPlane, refPlane, c, tap4, L, region, t2, lane );
      if( e < b.e ) { b.e = e; b.x = c.mvx; b.y = c.mvy; }
    }
    if( bx > 0 )
    {
      c.mvx = left.x; c.mvy = left.y;
      const int e = mctf_warp_error( orgPlane, refPlane, c, tap4, L, region, t2, lane );
      if( e < b.e ) { b.e = e; b.x = c.mvx; b.y = c.mvy; }
    }
    left = b;
    if( lane == 0 )
    {
      best[i].x = b.x; best[i].y = b.y; best[i].e = b.e;
      __threadfence();
      atomicExch( &progress[1 + by], bx + 1 );
    }
    __syncwarp();
  }
}

// stage E and the hand-over to the next level / the apply stage: field[by][bx] = { x, y, error, rmsme } for the blocks inside the out_w x out_h array
__global__ void mctf_final_kernel( MctfGeom g, const MctfBest* __restrict__ best, const double* __restrict__ var, int doubleRes, int bitDepth, vvb_mctf_mv* __restrict__ field )
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if( i >= g.n ) return;
  const int by = i / g.bxn, bx = i - by * g.bxn;
  if( bx >= g.outW || by >= g.outH ) return;
  int X, Y, W, H; mctf_block_of( g, i, X, Y, W, H );
  const MctfBest b = best[i];
  vvb_mctf_mv m; m.x = b.x; m.y = b.y; m.error = b.e; m.rmsme = 0xffff; m.pad = 0;
  if( doubleRes )
  {
    const double bdScale = (double)( 1 << ( 2 * ( 10 - bitDepth ) ) );
    const double wh = (double) W * (double) H;
    const double dvar = var[i] * bdScale;
    const double mse = (double) b.e * bdScale / wh;
    m.error = (int)( 20 * ( ( (double) b.e * bdScale + 5.0 ) / ( dvar + 5.0 ) ) + mse / 50.0 );
    m.rmsme = (uint16_t)(int)( 0.5 + sqrt( mse ) );
  }
  field[by * g.outW + bx] = m;
}

// MCTF::subsampleLuma (MCTF.cpp:1072-1097) with the border replication of PelStorage::extendBorderPel: dst(x, y) for x, y in [-margin, size + margin)
__global__ void mctf_subsample_kernel( const __grid_constant__ Plane src, int16_t* __restrict__ dstOrigin, int dstStride, int dw, int dh, int margin )
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x - margin, y = blockIdx.y * blockDim.y + threadIdx.y - margin;
  if( x >= dw + margin || y >= dh + margin ) return;
  const int cx = min( max( x, 0 ), dw - 1 ), cy = min( max( y, 0 ), dh - 1 );
  const int16_t* p = src.origin + (ptrdiff_t)( 2 * cy ) * src.stride + 2 * cx;
  dstOrigin[(ptrdiff_t) y * dstStride + x] = (int16_t)( ( (int) p[0] + (int) p[src.stride] + (int) p[1] + (int) p[src.stride + 1] + 2 ) >> 2 );
}

} // namespace vvb
