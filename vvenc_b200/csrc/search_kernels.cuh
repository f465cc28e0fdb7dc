// search_kernels.cuh -- integer motion-search SAD sweeps for sm_100a.
//
//  * sad_search_kernel : dense full search, InterSearch::xPatternSearch (EncoderLib/InterSearch.cpp:2209-2251).
//    One CTA per block; the (w + range) x (h + range) reference window and the original block are staged in shared
//    memory once, every thread then owns strips of 8 horizontally adjacent candidates and slides the original row over
//    a register-resident window row (packed 16x2 SAD: VIMNMX.S16x2 + IDP.2A), so each staged reference word is reused
//    for 8 candidates and each original word for 8 candidates x all strips.
//  * sad_pattern_kernel: the fixed TZ point pattern (xTZ8PointDiamondSearch / raster grid, InterSearch.cpp:557-758,
//    2491-2497) around a per-block start vector; candidates are read straight from the (L2-resident) reference plane.
//
// Both add the MV rate Distortion(sqrt(lambda)*bits) (CommonLib/RdCost.h:181-203) from a host-computed table and
// resolve the argmin with the reference's tie-break: first strictly smaller cost in evaluation order.
#pragma once
#include "common.cuh"
#include "dist_kernels.cuh"

namespace vvb {

struct MePar { int costScale, imvShift, subShift; MvCostTable tab; };

// tab: shared-memory copy of MePar::tab (dynamic indexing of the parameter bank would serialise per distinct address)
__device__ __forceinline__ uint32_t mv_cost( const MePar& p, const uint32_t* tab, int x, int y, int predHor, int predVer )
{
  const uint32_t bits = eg_bits( ( x * ( 1 << p.costScale ) - predHor ) >> p.imvShift ) + eg_bits( ( y * ( 1 << p.costScale ) - predVer ) >> p.imvShift );
  return tab[bits < VVB_MVCOST_ENTRIES ? bits : VVB_MVCOST_ENTRIES - 1];
}

// lexicographic (cost, order) minimum -> "first strictly smaller wins"
struct BestKey { unsigned long long cost; uint32_t order; };
__device__ __forceinline__ bool better( unsigned long long c, uint32_t o, unsigned long long bc, uint32_t bo ) { return c < bc || ( c == bc && o < bo ); }

// ---------------------------------------------------------------------------------------------------------------
// dense full search
// ---------------------------------------------------------------------------------------------------------------
#define SS_STRIP 8          // candidates per strip (consecutive dx)
#define SS_XCHUNK 16        // original pels consumed per inner step

// smem layout: win[(h + ny - 1)][ws] int16 (ws even, >= w + nxPad), org[h][w]
__global__ void __launch_bounds__( 256 ) sad_search_kernel( const __grid_constant__ Plane orgPlane, const __grid_constant__ Plane refPlane,
                                                            const vvb_block* __restrict__ blocks, int w, int h, const __grid_constant__ MePar par,
                                                            uint32_t* __restrict__ sadTables, int tableStride, vvb_best* __restrict__ bestOut )
{
  extern __shared__ __align__( 16 ) unsigned char smemRaw[];
  __shared__ uint32_t sMv[VVB_MVCOST_ENTRIES];
  for( int i = threadIdx.x; i < VVB_MVCOST_ENTRIES; i += blockDim.x ) sMv[i] = par.tab.cost[i];
  const vvb_block blk = blocks[blockIdx.x];
  const int nx = blk.right - blk.left + 1, ny = blk.bottom - blk.top + 1;
  const int nStrips = ( nx + SS_STRIP - 1 ) / SS_STRIP;
  const int winW = w + nStrips * SS_STRIP;                 // multiple of 8 pels beyond w: every strip can read w + 8 pels (+1 word slack below)
  const int ws = winW + 8;                                  // row pitch in pels (even, keeps 16-byte alignment of rows)
  const int winH = h + ny - 1;
  int16_t* win  = reinterpret_cast<int16_t*>( smemRaw );
  int16_t* orgS = win + (size_t) winH * ws;

  // ---- stage window (origin = block position + (left, top)) and original block
  {
    const int16_t* src = refPlane.origin + (ptrdiff_t)( blk.y + blk.top ) * refPlane.stride + blk.x + blk.left;
    const int validW = w + nx - 1;
    for( int i = threadIdx.x; i < winH * ws; i += blockDim.x )
    {
      const int r = i / ws, c = i - r * ws;
      win[i] = c < validW ? __ldg( src + (ptrdiff_t) r * refPlane.stride + c ) : (int16_t) 0;
    }
    const int16_t* so = orgPlane.origin + (ptrdiff_t) blk.y * orgPlane.stride + blk.x;
    for( int i = threadIdx.x; i < w * h; i += blockDim.x )
    {
      const int r = i / w, c = i - r * w;
      orgS[i] = __ldg( so + (ptrdiff_t) r * orgPlane.stride + c );
    }
  }
  __syncthreads();

  const int step = 1 << par.subShift;
  unsigned long long bestCost = ~0ull; uint32_t bestOrder = 0xffffffffu, bestSad = 0;

  const int items = ny * nStrips;
  for( int it = threadIdx.x; it < items; it += blockDim.x )
  {
    const int cy = it / nStrips, st = it - cy * nStrips;
    const int cx0 = st * SS_STRIP;
    int acc[SS_STRIP];
#pragma unroll
    for( int k = 0; k < SS_STRIP; k++ ) acc[k] = 0;

    for( int y = 0; y < h; y += step )
    {
      const uint32_t* orow = reinterpret_cast<const uint32_t*>( orgS + y * w );
      const uint32_t* rrow = reinterpret_cast<const uint32_t*>( win + ( cy + y ) * ws + cx0 );   // cx0 multiple of 8 -> 16-byte aligned
      if( w >= SS_XCHUNK )
      {
        for( int x = 0; x < w; x += SS_XCHUNK )
        {
          uint32_t o[SS_XCHUNK / 2], r[SS_XCHUNK / 2 + SS_STRIP / 2 + 1];
#pragma unroll
          for( int i = 0; i < SS_XCHUNK / 2; i += 4 ) *reinterpret_cast<uint4*>( &o[i] ) = *reinterpret_cast<const uint4*>( orow + x / 2 + i );
#pragma unroll
          for( int i = 0; i < SS_XCHUNK / 2 + SS_STRIP / 2; i += 4 ) *reinterpret_cast<uint4*>( &r[i] ) = *reinterpret_cast<const uint4*>( rrow + x / 2 + i );
          r[SS_XCHUNK / 2 + SS_STRIP / 2] = rrow[x / 2 + SS_XCHUNK / 2 + SS_STRIP / 2];
#pragma unroll
          for( int k = 0; k < SS_STRIP; k++ )
          {
#pragma unroll
            for( int i = 0; i < SS_XCHUNK / 2; i++ )
            {
              const uint32_t rv = ( k & 1 ) ? __funnelshift_r( r[i + k / 2], r[i + k / 2 + 1], 16 ) : r[i + k / 2];
              acc[k] = sad2_acc( o[i], rv, acc[k] );
            }
          }
        }
      }
      else
      {
        // w = 4 or 8 (2 or 4 words of original per row); w == 2 handled as a single word
        const int nw = w >> 1;
        uint32_t r[4 + SS_STRIP / 2 + 1];
#pragma unroll
        for( int i = 0; i < 4 + SS_STRIP / 2 + 1; i++ ) r[i] = ( i < nw + SS_STRIP / 2 + 1 ) ? rrow[i] : 0u;
#pragma unroll
        for( int i = 0; i < 4; i++ )
        {
          if( i < nw )
          {
            const uint32_t ov = orow[i];
#pragma unroll
            for( int k = 0; k < SS_STRIP; k++ )
            {
              const uint32_t rv = ( k & 1 ) ? __funnelshift_r( r[i + k / 2], r[i + k / 2 + 1], 16 ) : r[i + k / 2];
              acc[k] = sad2_acc( ov, rv, acc[k] );
            }
          }
        }
      }
    }

    const int dy = blk.top + cy;
#pragma unroll
    for( int k = 0; k < SS_STRIP; k++ )
    {
      const int cx = cx0 + k;
      if( cx < nx )
      {
        const uint32_t sad = (uint32_t) acc[k] << par.subShift;
        const uint32_t order = (uint32_t)( cy * nx + cx );
        if( sadTables ) sadTables[(size_t) blockIdx.x * tableStride + order] = sad;
        const unsigned long long c = (unsigned long long) sad + mv_cost( par, sMv, blk.left + cx, dy, blk.pred_hor, blk.pred_ver );
        if( better( c, order, bestCost, bestOrder ) ) { bestCost = c; bestOrder = order; bestSad = sad; }
      }
    }
  }

  // ---- block argmin (cost, raster order)
  __shared__ unsigned long long sCost[8];
  __shared__ uint32_t sOrder[8], sSad[8];
#pragma unroll
  for( int m = 16; m > 0; m >>= 1 )
  {
    const unsigned long long oc = __shfl_xor_sync( 0xffffffffu, bestCost, m );
    const uint32_t oo = __shfl_xor_sync( 0xffffffffu, bestOrder, m ), os = __shfl_xor_sync( 0xffffffffu, bestSad, m );
    if( better( oc, oo, bestCost, bestOrder ) ) { bestCost = oc; bestOrder = oo; bestSad = os; }
  }
  const int warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5;
  if( ( threadIdx.x & 31 ) == 0 ) { sCost[warp] = bestCost; sOrder[warp] = bestOrder; sSad[warp] = bestSad; }
  __syncthreads();
  if( threadIdx.x == 0 )
  {
    for( int i = 1; i < nWarps; i++ )
      if( better( sCost[i], sOrder[i], bestCost, bestOrder ) ) { bestCost = sCost[i]; bestOrder = sOrder[i]; bestSad = sSad[i]; }
    vvb_best b;
    const int cy = bestOrder / nx, cx = bestOrder - cy * nx;
    b.dx = (int16_t)( blk.left + cx ); b.dy = (int16_t)( blk.top + cy ); b.sad = bestSad; b.cost = bestCost;
    bestOut[blockIdx.x] = b;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// fixed pattern around a start vector; one CTA per block, G-lane groups take candidates round-robin
// ---------------------------------------------------------------------------------------------------------------
template<int G>
__global__ void __launch_bounds__( 128 ) cost_pattern_kernel( const __grid_constant__ Plane orgPlane, const __grid_constant__ Plane refPlane,
                                                              const vvb_block* __restrict__ blocks, int w, int h, int fam, const vvb_mv* __restrict__ pattern, int K,
                                                              const __grid_constant__ MePar par, uint32_t* __restrict__ sadOut, vvb_best* __restrict__ bestOut )
{
  __shared__ uint32_t sMv[VVB_MVCOST_ENTRIES];
  for( int i = threadIdx.x; i < VVB_MVCOST_ENTRIES; i += blockDim.x ) sMv[i] = par.tab.cost[i];
  __syncthreads();
  const vvb_block blk = blocks[blockIdx.x];
  const int lg = threadIdx.x & ( G - 1 );
  const int group = threadIdx.x / G, nGroups = blockDim.x / G;
  const int16_t* org = orgPlane.origin + (ptrdiff_t) blk.y * orgPlane.stride + blk.x;
  unsigned long long bestCost = ~0ull; uint32_t bestOrder = 0xffffffffu, bestSad = 0;

  for( int k = group; k < K; k += nGroups )
  {
    const vvb_mv pm = pattern[k];
    const int mx = blk.start_x + pm.dx, my = blk.start_y + pm.dy;
    const bool inside = mx >= blk.left && mx <= blk.right && my >= blk.top && my <= blk.bottom;      // SearchRange clip, InterSearch.cpp:576-620
    uint32_t sad = 0xffffffffu;
    if( inside )        // uniform per group
    {
      const int16_t* cur = refPlane.origin + (ptrdiff_t)( blk.y + my ) * refPlane.stride + blk.x + mx;
      sad = (uint32_t) group_dist<G>( fam, org, orgPlane.stride, cur, refPlane.stride, w, h, par.subShift, lg );
      const unsigned long long c = (unsigned long long) sad + mv_cost( par, sMv, mx, my, blk.pred_hor, blk.pred_ver );
      if( better( c, (uint32_t) k, bestCost, bestOrder ) ) { bestCost = c; bestOrder = (uint32_t) k; bestSad = sad; }
    }
    if( sadOut && lg == 0 ) sadOut[(size_t) blockIdx.x * K + k] = sad;
  }
  if( !bestOut ) return;
  __shared__ unsigned long long sCost[32];
  __shared__ uint32_t sOrder[32], sSad[32];
  if( lg == 0 ) { sCost[group] = bestCost; sOrder[group] = bestOrder; sSad[group] = bestSad; }
  __syncthreads();
  if( threadIdx.x == 0 )
  {
    for( int i = 1; i < nGroups; i++ )
      if( better( sCost[i], sOrder[i], bestCost, bestOrder ) ) { bestCost = sCost[i]; bestOrder = sOrder[i]; bestSad = sSad[i]; }
    vvb_best b;
    if( bestOrder == 0xffffffffu ) { b.dx = 0; b.dy = 0; b.sad = 0xffffffffu; b.cost = ~0ull; }
    else { const vvb_mv pm = pattern[bestOrder]; b.dx = (int16_t)( blk.start_x + pm.dx ); b.dy = (int16_t)( blk.start_y + pm.dy ); b.sad = bestSad; b.cost = bestCost; }
    bestOut[blockIdx.x] = b;
  }
}

// chains device-resident stages: the best vector of a search becomes the start / prediction offset of the next stage
__global__ void blocks_set_start_kernel( vvb_block* __restrict__ blocks, const vvb_best* __restrict__ best, int n )
{
  for( int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x )
  {
    blocks[i].start_x = best[i].dx; blocks[i].start_y = best[i].dy;
  }
}

} // namespace vvb
