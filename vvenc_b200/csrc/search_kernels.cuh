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

// smem layout: win[winH][ws] int16 | org[h][w] int16 | box[winH][nxp] uint32 | bitsX[nxp], bitsY[ny] uint8 (packed in uint32 words)
//
// SAD via  sum|a-b| = sum a + sum b - 2 sum min(a,b):
//   sum a           : once per block
//   sum b (box sum) : for every candidate from row-sliding sums of the staged window (O(window) work, shared by all candidates)
//   sum min(a,b)    : the only per-(candidate, pel) work: VIMNMX.S16x2 + IDP.2A per pel PAIR, i.e. one instruction per pel difference
// All three are exact integers, so the result is bit-identical to the direct sum.
struct SearchSmem { int ws, winH, nxp, offOrg, offBox, offBits, total; };

__host__ __device__ inline SearchSmem search_smem( int w, int h, int nx, int ny )
{
  SearchSmem s;
  const int nStrips = ( nx + SS_STRIP - 1 ) / SS_STRIP;
  s.nxp  = nStrips * SS_STRIP;
  s.ws   = w + s.nxp + 8;
  s.winH = h + ny - 1;
  s.offOrg  = s.winH * s.ws * 2;                       // bytes
  s.offBox  = s.offOrg + w * h * 2;
  s.offBits = s.offBox + s.winH * s.nxp * 4;
  s.total   = s.offBits + ( ( s.nxp + ny + 15 ) & ~15 ) * 4;
  return s;
}

__global__ void __launch_bounds__( 384 ) sad_search_kernel( const __grid_constant__ Plane orgPlane, const __grid_constant__ Plane refPlane,
                                                            const vvb_block* __restrict__ blocks, int w, int h, const __grid_constant__ MePar par,
                                                            uint32_t* __restrict__ sadTables, int tableStride, vvb_best* __restrict__ bestOut )
{
  extern __shared__ __align__( 16 ) unsigned char smemRaw[];
  __shared__ uint32_t sMv[VVB_MVCOST_ENTRIES];
  __shared__ int sSumA;
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nWarps = nthr >> 5;
  for( int i = tid; i < VVB_MVCOST_ENTRIES; i += nthr ) sMv[i] = par.tab.cost[i];
  if( tid == 0 ) sSumA = 0;
  const vvb_block blk = blocks[blockIdx.x];
  const int nx = blk.right - blk.left + 1, ny = blk.bottom - blk.top + 1;
  const int nStrips = ( nx + SS_STRIP - 1 ) / SS_STRIP;
  const SearchSmem L = search_smem( w, h, nx, ny );
  const int ws = L.ws, winH = L.winH, nxp = L.nxp;
  int16_t*  win   = reinterpret_cast<int16_t*>( smemRaw );
  int16_t*  orgS  = reinterpret_cast<int16_t*>( smemRaw + L.offOrg );
  uint32_t* box   = reinterpret_cast<uint32_t*>( smemRaw + L.offBox );      // [winH][nxp]: row sums first, then (in place) box sums for rows < ny
  int*      bitsX = reinterpret_cast<int*>( smemRaw + L.offBits );          // [nxp]
  int*      bitsY = bitsX + nxp;                                            // [ny]
  const int step = 1 << par.subShift;

  // ---- stage window (origin = block position + (left, top)) and original block; one window row per warp pass
  {
    const int16_t* src = refPlane.origin + (ptrdiff_t)( blk.y + blk.top ) * refPlane.stride + blk.x + blk.left;
    const int validW = w + nx - 1;
    // 32-bit words when the window start is even (plane rows are even-pitched), 16 independent loads in flight per thread:
    // the staging loop is latency bound otherwise (one L2 round trip per iteration).
    const bool even = ( ( (uintptr_t) src & 3 ) == 0 ) && ( ( refPlane.stride & 1 ) == 0 );
    if( even )
    {
      const int wpr = ws >> 1, total = winH * wpr, validWords = ( validW + 1 ) >> 1;
      uint32_t* win32 = reinterpret_cast<uint32_t*>( win );
      for( int i0 = tid; i0 < total; i0 += nthr * 16 )
      {
        uint32_t v[16];
#pragma unroll
        for( int u = 0; u < 16; u++ )
        {
          const int i = i0 + u * nthr;
          v[u] = 0u;
          if( i < total )
          {
            const int r = i / wpr, c = i - r * wpr;
            if( c < validWords ) v[u] = __ldg( reinterpret_cast<const uint32_t*>( src + (ptrdiff_t) r * refPlane.stride ) + c );
          }
        }
#pragma unroll
        for( int u = 0; u < 16; u++ ) { const int i = i0 + u * nthr; if( i < total ) win32[i] = v[u]; }
      }
    }
    else
    {
      const int total = winH * ws;
      for( int i0 = tid; i0 < total; i0 += nthr * 16 )
      {
        int16_t v[16];
#pragma unroll
        for( int u = 0; u < 16; u++ )
        {
          const int i = i0 + u * nthr;
          v[u] = 0;
          if( i < total )
          {
            const int r = i / ws, c = i - r * ws;
            if( c < validW ) v[u] = __ldg( src + (ptrdiff_t) r * refPlane.stride + c );
          }
        }
#pragma unroll
        for( int u = 0; u < 16; u++ ) { const int i = i0 + u * nthr; if( i < total ) win[i] = v[u]; }
      }
    }
    const int16_t* so = orgPlane.origin + (ptrdiff_t) blk.y * orgPlane.stride + blk.x;
    int sumA = 0;
    for( int i = tid; i < w * h; i += nthr )
    {
      const int r = i / w, c = i - r * w;
      const int16_t v = __ldg( so + (ptrdiff_t) r * orgPlane.stride + c );
      orgS[i] = v;
      if( ( r & ( step - 1 ) ) == 0 ) sumA += v;
    }
#pragma unroll
    for( int m = 16; m > 0; m >>= 1 ) sumA += __shfl_xor_sync( 0xffffffffu, sumA, m );
    __syncthreads();                                   // sSumA initialised, window visible
    if( lane == 0 && sumA ) atomicAdd( &sSumA, sumA );
    // MV-rate bit counts per column / row of the window (RdCost.h:183-203)
    for( int i = tid; i < nxp; i += nthr ) bitsX[i] = (int) eg_bits( ( ( blk.left + i ) * ( 1 << par.costScale ) - blk.pred_hor ) >> par.imvShift );
    for( int i = tid; i < ny;  i += nthr ) bitsY[i] = (int) eg_bits( ( ( blk.top  + i ) * ( 1 << par.costScale ) - blk.pred_ver ) >> par.imvShift );
  }
  // ---- row sums Hs[r][cx] = sum_{x<w} win[r][cx+x]: one task = (row, strip of 8 cx)
  for( int t = tid; t < winH * nStrips; t += nthr )
  {
    const int r = t / nStrips, st = t - r * nStrips;
    const int16_t* row = win + r * ws + st * SS_STRIP;
    int s = 0;
    for( int x = 0; x < w; x++ ) s += row[x];
    uint32_t* dst = box + r * nxp + st * SS_STRIP;
    dst[0] = (uint32_t) s;
#pragma unroll
    for( int k = 1; k < SS_STRIP; k++ ) { s += row[w + k - 1] - row[k - 1]; dst[k] = (uint32_t) s; }
  }
  __syncthreads();
  // ---- box sums in place: B[cy][cx] = sum_{k < h/step} Hs[cy + k*step][cx]; rows of one phase (cy mod step) only depend on that phase
  {
    const int m = h >> par.subShift;
    for( int t = tid; t < nxp * step; t += nthr )
    {
      const int cx = t % nxp, p = t / nxp;
      if( p < ny )
      {
        int cur = 0;
        for( int k = 0; k < m; k++ ) cur += (int) box[( p + k * step ) * nxp + cx];
        int prevTop = (int) box[p * nxp + cx];
        box[p * nxp + cx] = (uint32_t) cur;
        for( int cy = p + step; cy < ny; cy += step )
        {
          cur += (int) box[( cy + ( m - 1 ) * step ) * nxp + cx] - prevTop;
          prevTop = (int) box[cy * nxp + cx];
          box[cy * nxp + cx] = (uint32_t) cur;
        }
      }
    }
  }
  __syncthreads();
  const int sumA = sSumA;

  unsigned long long bestCost = ~0ull; uint32_t bestOrder = 0xffffffffu, bestSad = 0;
  const int items = ny * nStrips;
  for( int it = tid; it < items; it += nthr )
  {
    const int cy = it / nStrips, st = it - cy * nStrips;
    const int cx0 = st * SS_STRIP;
    int acc[SS_STRIP];                                 // = - sum min(org, ref)
#pragma unroll
    for( int k = 0; k < SS_STRIP; k++ ) acc[k] = 0;

    for( int y = 0; y < h; y += step )
    {
      const uint32_t* orow = reinterpret_cast<const uint32_t*>( orgS + y * w );
      const uint32_t* rrow = reinterpret_cast<const uint32_t*>( win + ( cy + y ) * ws + cx0 );   // cx0 multiple of 8 -> 16-byte aligned when ws % 8 == 0
      if( w >= SS_XCHUNK )
      {
        for( int x = 0; x < w; x += SS_XCHUNK )
        {
          uint32_t o[SS_XCHUNK / 2], r[SS_XCHUNK / 2 + SS_STRIP / 2];
#pragma unroll
          for( int i = 0; i < SS_XCHUNK / 2; i += 4 ) *reinterpret_cast<uint4*>( &o[i] ) = *reinterpret_cast<const uint4*>( orow + x / 2 + i );
#pragma unroll
          for( int i = 0; i < SS_XCHUNK / 2 + SS_STRIP / 2; i += 4 ) *reinterpret_cast<uint4*>( &r[i] ) = *reinterpret_cast<const uint4*>( rrow + x / 2 + i );
#pragma unroll
          for( int k = 0; k < SS_STRIP; k++ )
          {
#pragma unroll
            for( int i = 0; i < SS_XCHUNK / 2; i++ )
            {
              const uint32_t rv = ( k & 1 ) ? __funnelshift_r( r[i + k / 2], r[i + k / 2 + 1], 16 ) : r[i + k / 2];
              acc[k] = __dp2a_lo( (int) __vmins2( o[i], rv ), (int) 0x0000ffffu, acc[k] );
            }
          }
        }
      }
      else
      {
        // w = 4 or 8: 2 or 4 words of original per row
        const int nw = w >> 1;
        uint32_t r[4 + SS_STRIP / 2];
#pragma unroll
        for( int i = 0; i < 4 + SS_STRIP / 2; i++ ) r[i] = ( i < nw + SS_STRIP / 2 ) ? rrow[i] : 0u;
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
              acc[k] = __dp2a_lo( (int) __vmins2( ov, rv ), (int) 0x0000ffffu, acc[k] );
            }
          }
        }
      }
    }

    const int by = bitsY[cy];
    const uint32_t* brow = box + cy * nxp + cx0;
#pragma unroll
    for( int k = 0; k < SS_STRIP; k++ )
    {
      const int cx = cx0 + k;
      if( cx < nx )
      {
        const uint32_t sad = (uint32_t)( sumA + (int) brow[k] + 2 * acc[k] ) << par.subShift;
        const uint32_t order = (uint32_t)( cy * nx + cx );
        if( sadTables ) sadTables[(size_t) blockIdx.x * tableStride + order] = sad;
        const uint32_t bits = (uint32_t)( bitsX[cx] + by );
        const unsigned long long c = (unsigned long long) sad + sMv[bits < VVB_MVCOST_ENTRIES ? bits : VVB_MVCOST_ENTRIES - 1];
        if( better( c, order, bestCost, bestOrder ) ) { bestCost = c; bestOrder = order; bestSad = sad; }
      }
    }
  }

  // ---- block argmin (cost, raster order)
  __shared__ unsigned long long sCost[12];
  __shared__ uint32_t sOrder[12], sSad[12];
#pragma unroll
  for( int m = 16; m > 0; m >>= 1 )
  {
    const unsigned long long oc = __shfl_xor_sync( 0xffffffffu, bestCost, m );
    const uint32_t oo = __shfl_xor_sync( 0xffffffffu, bestOrder, m ), os = __shfl_xor_sync( 0xffffffffu, bestSad, m );
    if( better( oc, oo, bestCost, bestOrder ) ) { bestCost = oc; bestOrder = oo; bestSad = os; }
  }
  if( lane == 0 ) { sCost[warp] = bestCost; sOrder[warp] = bestOrder; sSad[warp] = bestSad; }
  __syncthreads();
  if( tid == 0 )
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
