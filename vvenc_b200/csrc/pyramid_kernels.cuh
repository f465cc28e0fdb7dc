// pyramid_kernels.cuh -- SAD pyramid of the dense integer search with every level formed inside one CTA (sm_100a).
//
// Replaces, for 8x8 base blocks without row sub-sampling, the pair sad_search_kernel<.., PARENT> + sad_table_sum_kernel: one CTA owns one ROOT block of
// 16x16, 32x32 or 64x64 pels (LV = 2, 3, 4 levels) together with all of its quad-tree descendants, stages the reference window of the whole root once and
// produces the InterSearch::xPatternSearch result (EncoderLib/InterSearch.cpp:2209-2251: every vector of the range, MV rate of the block's own predictor,
// first strictly smaller cost in raster order) of EVERY block of every level.  Pel work happens once, at the 8x8 level; a 16x16 cost is the sum of its
// four children's SADs in registers; 32x32 tables are accumulated in shared memory and the 64x64 table is their sum -- no cost table ever leaves the SM
// (the previous design moved about 1.2 GB of 16x16 / 32x32 tables through HBM per 2160p picture).
//
// Arithmetic per (8x8 block, candidate):  SAD = sum a + sum b - 2 sum min(a,b)
//   sum a : per block, once;  sum b : 8x8 box sums of the window (uint16 table V);  sum min : VIMNMX.S16x2 (alu pipe) + IDP.2A (fma pipe) per pel pair.
// Both pipes issue every other cycle per scheduler (B300_MICROARCH.md "fma vs alu split"), so one min + one dot product per pel pair is the floor of this
// formulation; everything else is kept off the alu pipe where possible:
//   * odd-offset candidates read a second, one-pel-shifted copy of the window (no funnel shifts),
//   * a thread evaluates TWO vertically adjacent candidate rows for a strip of 8 vectors: the nine window rows they need are loaded once (LDS.128) and
//     every original row serves both (shared-memory traffic per candidate halves),
//   * the epilogue per candidate is IDP.4A (MV bits of column + row -> table offset), LDS (rate table pre-multiplied by 8, one table per strip slot so
//     that the slot index is part of the entry), IDP.2A (unpack the box sum and add it), IMAD (parent sum), IMAD (key) on the fma pipe and one VIMNMX.
//   * strip slots past the end of the range carry MV-bit count 250 and read rate-table entries that can never win; no predicate per candidate.
#pragma once
#include "search_kernels.cuh"

namespace vvb {

#define PYR_MAX_THREADS 640
#define PYR_MVN         296                      // rate-table entries per strip slot: 0..79 real, the rest "never wins" (padded columns index 250 + row bits)
#define PYR_PAD_BITS    250
#define PYR_NEVER       ( 1u << 26 )             // cost no real candidate reaches; 4 * PYR_NEVER * 8 still fits 32 bits

struct PyrLevels { const vvb_block* blocks[4]; vvb_best* best[4]; };

struct PyrSmem
{
  int nStrips, nxp, nyp, bStride, ws, winH, winWords, vRows, vPitch, nT, tStride;
  int offWin0, offWin1, offV, offOrg, offBits, offPred, offSumA, offKey32, offKey64, offMv8, offMvRaw, offT, total;   // bytes
};

template<int LV>
__host__ __device__ inline PyrSmem pyr_smem( int nx, int ny )
{
  constexpr int R = 8 << ( LV - 1 ), NB0 = 1 << ( 2 * ( LV - 1 ) ), NBLK = ( 4 * NB0 - 1 ) / 3;
  PyrSmem s;
  s.nStrips = ( nx + 7 ) >> 3;
  s.nxp     = s.nStrips * 8;
  s.nyp     = ( ny + 1 + 7 ) & ~7;                       // row B of the last pair may be one past the range
  s.bStride = s.nxp + s.nyp;                             // per block: column bits [nxp] (raw), row bits [nyp] (times 4); multiple of 8
  // row pitch: multiple of 8 pels (16-byte rows); pitch/8 == nStrips (mod 8) makes a warp's LDS.128 walk consecutive 16-byte chunks across rows
  int ws = R + s.nxp;
  const int want = ( ( s.nStrips - ( ws >> 3 ) ) % 8 + 8 ) % 8;
  if( want <= 2 ) ws += 8 * want;
  s.ws      = ws;
  s.winH    = R + ny - 1;
  s.winWords = ( s.winH * s.ws + 16 ) >> 1;              // + overrun for the shifted copy
  s.vRows   = s.winH - 7;
  s.vPitch  = R - 8 + s.nxp;
  s.nT      = LV == 4 ? 4 : ( LV == 3 ? 1 : 0 );
  s.tStride = ny * s.nxp;
  // fixed-size tables first: their shared-memory addresses are then link-time constants (the rate look-up becomes LDS [reg + imm])
  int o = 0;
  s.offMv8  = o;   o += 8 * PYR_MVN * 4;
  s.offMvRaw = o;  o += VVB_MVCOST_ENTRIES * 4;
  s.offKey64 = o;  o += 8 * 8;
  s.offKey32 = o;  o += ( ( NB0 + NB0 / 4 ) * 4 + 15 ) & ~15;
  s.offSumA = o;   o += ( NB0 * 4 + 15 ) & ~15;
  s.offPred = o;   o += ( NBLK * 8 + 15 ) & ~15;
  s.offWin0 = o;   o += s.winWords * 4;
  s.offWin1 = o;   o += s.winWords * 4;
  s.offV    = o;   o += ( ( s.vRows * s.vPitch * 2 ) + 15 ) & ~15;
  s.offOrg  = o;   o += R * R * 2;
  s.offBits = o;   o += NBLK * s.bStride;
  s.offT    = ( o + 15 ) & ~15;
  const int tBytes = s.nT * s.tStride * 4, hsBytes = s.winH * s.vPitch * 2;      // the row-sum scratch of the prologue lives where the tables go later
  s.total   = s.offT + ( tBytes > hsBytes ? tBytes : hsBytes ) + 16;
  return s;
}

__device__ __forceinline__ int pyr_compact( int v ) { v &= 0x55555555; v = ( v | ( v >> 1 ) ) & 0x33333333; v = ( v | ( v >> 2 ) ) & 0x0f0f0f0f; return ( v | ( v >> 4 ) ) & 0xff; }

// one candidate row of one member: box sum + MV rate -> 8 packed (cost * 8 + slot) keys, running minimum; the SAD goes into the parent's sum
__device__ __forceinline__ uint32_t pyr_finish_row( const int (&acc)[8], const uint16_t* __restrict__ vrow, uint2 bw, uint32_t by4, const unsigned char* __restrict__ mv8,
                                                    uint32_t one, uint32_t eight, uint32_t (&ps)[8] )
{
  const uint4 vw = *reinterpret_cast<const uint4*>( vrow );
  const uint32_t v[4] = { vw.x, vw.y, vw.z, vw.w };
  uint32_t bk = 0xffffffffu;
#pragma unroll
  for( int k = 0; k < 8; k++ )
  {
    const uint32_t idx4 = __dp4a( k < 4 ? bw.x : bw.y, 4u << ( 8 * ( k & 3 ) ), by4 );                  // 4 * (column bits + row bits)
    const uint32_t mvk  = *reinterpret_cast<const uint32_t*>( mv8 + k * ( PYR_MVN * 4 ) + idx4 );        // rate * 8 + k
    const uint32_t sad  = __dp2a_lo( v[k >> 1], ( k & 1 ) ? 0x0100u : 0x0001u, (uint32_t) acc[k] );      // box sum + (sum a - 2 sum min)
    ps[k] = sad * one + ps[k];                                                                             // IMAD: keeps the add off the alu pipe
    const uint32_t key = sad * eight + mvk;
    bk = min( bk, key );
  }
  return bk;
}

template<int LV>
__global__ void __launch_bounds__( PYR_MAX_THREADS, 1 ) sad_pyramid8_kernel( const __grid_constant__ Plane orgPlane, const __grid_constant__ Plane refPlane,
                                                                             const __grid_constant__ PyrLevels lv, int rootFirst, int nx, int ny,
                                                                             const __grid_constant__ MePar par, uint32_t one, uint32_t eight )
{
  constexpr int R = 8 << ( LV - 1 ), NB0 = 1 << ( 2 * ( LV - 1 ) ), NQ = NB0 / 4, NBLK = ( 4 * NB0 - 1 ) / 3, LTOP = LV - 1;
  constexpr int OFF1 = NB0, OFF2 = NB0 + NQ, OFF3 = NB0 + NQ + NQ / 4;
  extern __shared__ __align__( 128 ) unsigned char smemRaw[];
  const PyrSmem L = pyr_smem<LV>( nx, ny );
  uint32_t* win0w = reinterpret_cast<uint32_t*>( smemRaw + L.offWin0 );
  uint32_t* win1w = reinterpret_cast<uint32_t*>( smemRaw + L.offWin1 );
  uint16_t* V     = reinterpret_cast<uint16_t*>( smemRaw + L.offV );
  int16_t*  orgS  = reinterpret_cast<int16_t*>( smemRaw + L.offOrg );
  unsigned char* bitsS = smemRaw + L.offBits;
  int2*     sPred = reinterpret_cast<int2*>( smemRaw + L.offPred );
  int*      sSumA = reinterpret_cast<int*>( smemRaw + L.offSumA );
  uint32_t* sKey32 = reinterpret_cast<uint32_t*>( smemRaw + L.offKey32 );
  unsigned long long* sKey64 = reinterpret_cast<unsigned long long*>( smemRaw + L.offKey64 );
  unsigned char* sMv8 = smemRaw + L.offMv8;
  uint32_t* sMvRaw = reinterpret_cast<uint32_t*>( smemRaw + L.offMvRaw );
  uint32_t* T     = reinterpret_cast<uint32_t*>( smemRaw + L.offT );
  uint16_t* Hs    = reinterpret_cast<uint16_t*>( smemRaw + L.offT );

  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31;
  const int root = rootFirst + blockIdx.x;
  const vvb_block rb = lv.blocks[LTOP][root];
  const int nxp = L.nxp, nStrips = L.nStrips, ws = L.ws, wsw = ws >> 1, winH = L.winH;
  const int ob = par.orderBits;

  // ---- the root's descendants: positions must be the z-order tiling of the root, ranges must equal the launch's range
  int geomOk = ( rb.right - rb.left + 1 == nx ) && ( rb.bottom - rb.top + 1 == ny );
  for( int t = tid; t < NBLK; t += nthr )
  {
    const int l = t < OFF1 ? 0 : ( t < OFF2 ? 1 : ( t < OFF3 ? 2 : 3 ) );
    const int i = t - ( l == 0 ? 0 : ( l == 1 ? OFF1 : ( l == 2 ? OFF2 : OFF3 ) ) );
    const vvb_block b = lv.blocks[l][( (size_t) root << ( 2 * ( LTOP - l ) ) ) + i];
    const int s = 8 << l;
    geomOk &= ( b.x == rb.x + pyr_compact( i ) * s ) && ( b.y == rb.y + pyr_compact( i >> 1 ) * s ) &&
              ( b.left == rb.left ) && ( b.right == rb.right ) && ( b.top == rb.top ) && ( b.bottom == rb.bottom );
    sPred[t] = make_int2( b.pred_hor, b.pred_ver );
  }
  geomOk = __syncthreads_and( geomOk );
  if( !geomOk )
  {
    // not a proper quad tree (or a block with another range): everything below this root is reported invalid, as the header promises
    for( int t = tid; t < NBLK; t += nthr )
    {
      const int l = t < OFF1 ? 0 : ( t < OFF2 ? 1 : ( t < OFF3 ? 2 : 3 ) );
      const int i = t - ( l == 0 ? 0 : ( l == 1 ? OFF1 : ( l == 2 ? OFF2 : OFF3 ) ) );
      vvb_best b; b.dx = 0; b.dy = 0; b.sad = 0xffffffffu; b.cost = ~0ull;
      lv.best[l][( (size_t) root << ( 2 * ( LTOP - l ) ) ) + i] = b;
    }
    return;
  }

  // ---- stage the window (zero beyond the valid columns) and the original root block
  {
    const int16_t* src = refPlane.origin + (ptrdiff_t)( rb.y + rb.top ) * refPlane.stride + rb.x + rb.left;
    const int validW = R + nx - 1;
    if( ( ( (uintptr_t) src & 3 ) == 0 ) && ( ( refPlane.stride & 1 ) == 0 ) )
    {
      const int total = winH * wsw, validWords = ( validW + 1 ) >> 1;
      const float inv = 1.0f / (float) wsw;
      for( int i0 = tid; i0 < total; i0 += nthr * 8 )
      {
        uint32_t v[8];
#pragma unroll
        for( int u = 0; u < 8; u++ )
        {
          const int i = i0 + u * nthr;
          v[u] = 0u;
          if( i < total )
          {
            const int r = fast_div( i, inv ), c = i - r * wsw;
            if( c < validWords ) v[u] = __ldg( reinterpret_cast<const uint32_t*>( src + (ptrdiff_t) r * refPlane.stride ) + c );
          }
        }
#pragma unroll
        for( int u = 0; u < 8; u++ ) { const int i = i0 + u * nthr; if( i < total ) win0w[i] = v[u]; }
      }
    }
    else
    {
      int16_t* win0 = reinterpret_cast<int16_t*>( win0w );
      const int total = winH * ws;
      const float inv = 1.0f / (float) ws;
      for( int i0 = tid; i0 < total; i0 += nthr * 8 )
      {
        int16_t v[8];
#pragma unroll
        for( int u = 0; u < 8; u++ )
        {
          const int i = i0 + u * nthr;
          v[u] = 0;
          if( i < total )
          {
            const int r = fast_div( i, inv ), c = i - r * ws;
            if( c < validW ) v[u] = __ldg( src + (ptrdiff_t) r * refPlane.stride + c );
          }
        }
#pragma unroll
        for( int u = 0; u < 8; u++ ) { const int i = i0 + u * nthr; if( i < total ) win0[i] = v[u]; }
      }
    }
    if( tid < 8 ) win0w[winH * wsw + tid] = 0u;                                   // overrun words read by the shifted copy
    const int16_t* so = orgPlane.origin + (ptrdiff_t) rb.y * orgPlane.stride + rb.x;
    for( int i = tid; i < R * R; i += nthr )
    {
      const int r = i / R, c = i - r * R;
      orgS[i] = __ldg( so + (ptrdiff_t) r * orgPlane.stride + c );
    }
    for( int i = tid; i < VVB_MVCOST_ENTRIES; i += nthr ) sMvRaw[i] = par.tab.cost[i];
    for( int i = tid; i < 8 * PYR_MVN; i += nthr )
    {
      const int k = i / PYR_MVN, b = i - k * PYR_MVN;
      reinterpret_cast<uint32_t*>( sMv8 )[i] = ( b < VVB_MVCOST_ENTRIES ? par.tab.cost[b] : PYR_NEVER ) * 8u + (uint32_t) k;
    }
    for( int i = tid; i < NB0 + NQ; i += nthr ) sKey32[i] = 0xffffffffu;
    if( tid < 8 ) sKey64[tid] = ~0ull;
  }
  __syncthreads();

  // ---- shifted copy, per-block sum a, row sums Hs[r][c] = sum_{x<8} win[r][c+x], MV bit counts
  {
    for( int i = tid; i < winH * wsw; i += nthr ) win1w[i] = __funnelshift_r( win0w[i], win0w[i + 1], 16 );
    for( int b = tid; b < NB0; b += nthr )
    {
      const int bx = pyr_compact( b ), by = pyr_compact( b >> 1 );
      int s = 0;
      for( int y = 0; y < 8; y++ )
      {
        const uint4 o = *reinterpret_cast<const uint4*>( orgS + ( by * 8 + y ) * R + bx * 8 );
        s = __dp2a_lo( (int) o.x, 0x0101, s ); s = __dp2a_lo( (int) o.y, 0x0101, s ); s = __dp2a_lo( (int) o.z, 0x0101, s ); s = __dp2a_lo( (int) o.w, 0x0101, s );
      }
      sSumA[b] = s;
    }
    const int cStrips = L.vPitch >> 3, nTasks = winH * cStrips;
    const float inv = 1.0f / (float) cStrips;
    for( int t = tid; t < nTasks; t += nthr )
    {
      const int r = fast_div( t, inv ), st = t - r * cStrips;
      const uint32_t* row = win0w + r * wsw + st * 4;
      const uint4 a = *reinterpret_cast<const uint4*>( row ), b = *reinterpret_cast<const uint4*>( row + 4 );
      const uint32_t w[8] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w };
      int p[16];
#pragma unroll
      for( int i = 0; i < 8; i++ ) { p[2 * i] = (int)( w[i] & 0xffffu ); p[2 * i + 1] = (int)( w[i] >> 16 ); }
      int s = p[0] + p[1] + p[2] + p[3] + p[4] + p[5] + p[6] + p[7];
      uint32_t o[4];
#pragma unroll
      for( int k = 0; k < 8; k++ )
      {
        if( k ) s += p[k + 7] - p[k - 1];
        if( k & 1 ) o[k >> 1] |= (uint32_t) s << 16; else o[k >> 1] = (uint32_t) s;
      }
      *reinterpret_cast<uint4*>( Hs + r * L.vPitch + st * 8 ) = make_uint4( o[0], o[1], o[2], o[3] );
    }
    for( int t = tid; t < NBLK * L.bStride; t += nthr )
    {
      const int bid = t / L.bStride, e = t - bid * L.bStride;
      const int2 pr = sPred[bid];
      unsigned char v;
      if( e < nxp ) v = e < nx ? (unsigned char) eg_bits( ( ( rb.left + e ) * ( 1 << par.costScale ) - pr.x ) >> par.imvShift ) : (unsigned char) PYR_PAD_BITS;
      else          v = (unsigned char)( 4u * eg_bits( ( ( rb.top + ( e - nxp ) ) * ( 1 << par.costScale ) - pr.y ) >> par.imvShift ) );
      bitsS[t] = v;
    }
  }
  __syncthreads();
  // ---- box sums V[r][c] = sum_{y<8} Hs[r+y][c]  (uint16: 64 * 1023 fits); a thread slides down a chunk of rows of one column pair
  {
    const int cPairs = L.vPitch >> 1, chunk = 16, nChunks = ( L.vRows + chunk - 1 ) / chunk;
    const uint32_t* Hs32 = reinterpret_cast<const uint32_t*>( Hs );
    uint32_t* V32 = reinterpret_cast<uint32_t*>( V );
    for( int t = tid; t < cPairs * nChunks; t += nthr )
    {
      const int ch = t / cPairs, c = t - ch * cPairs;
      const int r0 = ch * chunk, r1 = min( L.vRows, r0 + chunk );
      uint32_t s = 0;                                                           // two uint16 lanes, no carry: each lane stays below 2^16
      for( int y = 0; y < 8; y++ ) s += Hs32[( r0 + y ) * cPairs + c];
      V32[r0 * cPairs + c] = s;
      for( int r = r0 + 1; r < r1; r++ ) { s += Hs32[( r + 7 ) * cPairs + c] - Hs32[( r - 1 ) * cPairs + c]; V32[r * cPairs + c] = s; }
    }
  }
  __syncthreads();
  if( LV >= 3 ) { for( int i = tid; i < L.nT * L.tStride; i += nthr ) T[i] = 0u; }
  __syncthreads();

  // ---- candidates: item = (quad of four 8x8 members, pair of candidate rows, strip of 8 vectors)
  {
    // Lane mapping.  A quarter warp's LDS.128 is one wavefront when its 8 lanes read 8 consecutive 16-byte chunks: main items are groups of 8 adjacent strips
    // of one row pair (it = ((q * nPairs + pr) * nMain + st), st fastest); the strips left over when the range is not a multiple of 64 vectors follow as
    // tail items with the row pair as the fast index.
    const int nFull = nx >> 3, nCols = nx & 7;                // full strips of 8 vectors; the nx % 8 columns left of them are column items (below)
    const int nPairs = ( ny + 1 ) >> 1, nMain = nFull & ~7, nTail = nFull - nMain;
    const int perQm = nPairs * nMain, itemsMain = NQ * perQm, perQt = nPairs * nTail, items = itemsMain + NQ * perQt;
    const float invPerQm = 1.0f / (float) max( 1, perQm ), invMain = 1.0f / (float) max( 1, nMain ), invPerQt = 1.0f / (float) max( 1, perQt ), invPairs = 1.0f / (float) nPairs;
    const uint32_t* org32 = reinterpret_cast<const uint32_t*>( orgS );
    for( int base = 0; base < items; base += nthr )
    {
      const int it = base + tid;
      const bool active = it < items;
      const unsigned mask = __ballot_sync( 0xffffffffu, active );
      if( !active ) continue;
      int q, pr, st;
      if( it < itemsMain ) { q = fast_div( it, invPerQm ); const int rem = it - q * perQm; pr = fast_div( rem, invMain ); st = rem - pr * nMain; }
      else { const int i2 = it - itemsMain; q = fast_div( i2, invPerQt ); const int rem = i2 - q * perQt; const int ts = fast_div( rem, invPairs ); pr = rem - ts * nPairs; st = nMain + ts; }
      const int cy = 2 * pr, cx0 = 8 * st;
      const bool validB = cy + 1 < ny;
      const int lead = __ffs( mask ) - 1;
      const bool uni = __all_sync( mask, q == __shfl_sync( mask, q, lead ) );
      const int qx = pyr_compact( q ), qy = pyr_compact( q >> 1 );
      uint32_t psA[8], psB[8];
#pragma unroll
      for( int k = 0; k < 8; k++ ) { psA[k] = 0u; psB[k] = 0u; }
#pragma unroll 1
      for( int m = 0; m < 4; m++ )
      {
        const int bx8 = 2 * qx + ( m & 1 ), by8 = 2 * qy + ( m >> 1 ), b0 = 4 * q + m;
        const int sumA = sSumA[b0];
        int accA[8], accB[8];
#pragma unroll
        for( int k = 0; k < 8; k++ ) { accA[k] = sumA; accB[k] = sumA; }
        const uint32_t* op = org32 + ( by8 * 8 ) * ( R / 2 ) + bx8 * 4;
        const int wofs = ( by8 * 8 + cy ) * wsw + bx8 * 4 + ( cx0 >> 1 );
        const uint32_t* w0 = win0w + wofs;
        const uint32_t* w1 = win1w + wofs;
        uint4 oPrev = make_uint4( 0, 0, 0, 0 );
#pragma unroll
        for( int y = 0; y < 9; y++ )
        {
          const uint4 e0 = *reinterpret_cast<const uint4*>( w0 + y * wsw ), e1 = *reinterpret_cast<const uint4*>( w0 + y * wsw + 4 );
          const uint4 d0 = *reinterpret_cast<const uint4*>( w1 + y * wsw ), d1 = *reinterpret_cast<const uint4*>( w1 + y * wsw + 4 );
          const uint32_t e[8] = { e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w };
          const uint32_t d[8] = { d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w };
          uint4 oCur = oPrev;
          if( y < 8 )
          {
            oCur = *reinterpret_cast<const uint4*>( op + y * ( R / 2 ) );
            const uint32_t o[4] = { oCur.x, oCur.y, oCur.z, oCur.w };
#pragma unroll
            for( int k = 0; k < 8; k++ )
#pragma unroll
              for( int i = 0; i < 4; i++ )
                accA[k] = __dp2a_lo( (int) __vmins2( o[i], ( k & 1 ) ? d[i + ( k >> 1 )] : e[i + ( k >> 1 )] ), (int) 0x0000fefeu, accA[k] );
          }
          if( y > 0 )
          {
            const uint32_t o[4] = { oPrev.x, oPrev.y, oPrev.z, oPrev.w };
#pragma unroll
            for( int k = 0; k < 8; k++ )
#pragma unroll
              for( int i = 0; i < 4; i++ )
                accB[k] = __dp2a_lo( (int) __vmins2( o[i], ( k & 1 ) ? d[i + ( k >> 1 )] : e[i + ( k >> 1 )] ), (int) 0x0000fefeu, accB[k] );
          }
          oPrev = oCur;
        }
        // member epilogue
        const unsigned char* bb = bitsS + b0 * L.bStride;
        const uint2 bw = *reinterpret_cast<const uint2*>( bb + cx0 );
        const uint16_t* vrow = V + ( by8 * 8 + cy ) * L.vPitch + bx8 * 8 + cx0;
        uint32_t bk = pyr_finish_row( accA, vrow, bw, bb[nxp + cy], sMv8, one, eight, psA );
        uint32_t key = ( ( bk >> 3 ) << ob ) + (uint32_t)( cy * nx + cx0 ) + ( bk & 7u );
        if( validB )
        {
          bk = pyr_finish_row( accB, vrow + L.vPitch, bw, bb[nxp + cy + 1], sMv8, one, eight, psB );
          key = min( key, ( ( bk >> 3 ) << ob ) + (uint32_t)( ( cy + 1 ) * nx + cx0 ) + ( bk & 7u ) );
        }
        if( uni ) { key = __reduce_min_sync( mask, key ); if( lane == lead ) atomicMin( &sKey32[b0], key ); }
        else atomicMin( &sKey32[b0], key );
      }
      // the 16x16 parent of the quad: its SAD at a vector is the sum of the members' SADs
      {
        const unsigned char* bb = bitsS + ( OFF1 + q ) * L.bStride;
        const uint2 bw = *reinterpret_cast<const uint2*>( bb + cx0 );
        uint32_t* trow = LV >= 3 ? T + ( LV == 4 ? ( q >> 2 ) : 0 ) * L.tStride + cy * nxp + st : nullptr;      // table layout [cy][slot k][strip]: a warp's atomics spread over the banks
        uint32_t key = 0xffffffffu;
#pragma unroll
        for( int rowB = 0; rowB < 2; rowB++ )
        {
          if( rowB && !validB ) break;
          const uint32_t by4 = bb[nxp + cy + rowB];
          uint32_t bk = 0xffffffffu;
#pragma unroll
          for( int k = 0; k < 8; k++ )
          {
            const uint32_t ps = rowB ? psB[k] : psA[k];
            const uint32_t idx4 = __dp4a( k < 4 ? bw.x : bw.y, 4u << ( 8 * ( k & 3 ) ), by4 );
            const uint32_t mvk  = *reinterpret_cast<const uint32_t*>( sMv8 + k * ( PYR_MVN * 4 ) + idx4 );
            bk = min( bk, ps * eight + mvk );
            if( LV >= 3 ) atomicAdd( trow + rowB * nxp + k * nStrips, ps );
          }
          key = min( key, ( ( bk >> 3 ) << ob ) + (uint32_t)( ( cy + rowB ) * nx + cx0 ) + ( bk & 7u ) );
        }
        if( uni ) { key = __reduce_min_sync( mask, key ); if( lane == lead ) atomicMin( &sKey32[OFF1 + q], key ); }
        else atomicMin( &sKey32[OFF1 + q], key );
      }
    }

    // ---- the nx % 8 rightmost columns (one column for every +-R range): item = (quad, column, group of 8 vertically adjacent vectors).  A strip item would
    // spend a full strip of work on them (11 % of the kernel for 65 columns); here the thread keeps the member's eight original rows in registers and walks
    // the 15 window rows its 8 vectors touch: window row r meets original row r - c for vector c.
    if( nCols )
    {
      const int nV = ( ny + 7 ) >> 3, perQc = nCols * nV, itemsC = NQ * perQc;
      const float invPerQc = 1.0f / (float) perQc, invNv = 1.0f / (float) nV;
      for( int base = 0; base < itemsC; base += nthr )
      {
        const int it = base + tid;
        const bool active = it < itemsC;
        const unsigned mask = __ballot_sync( 0xffffffffu, active );
        if( !active ) continue;
        const int q = fast_div( it, invPerQc ), rem = it - q * perQc;
        const int ci = fast_div( rem, invNv ), g = rem - ci * nV;
        const int cx = 8 * nFull + ci, cy0 = 8 * g;
        const int lead = __ffs( mask ) - 1;
        const bool uni = __all_sync( mask, q == __shfl_sync( mask, q, lead ) );
        const int qx = pyr_compact( q ), qy = pyr_compact( q >> 1 );
        const uint32_t* wsrc = ( cx & 1 ) ? win1w : win0w;      // odd columns read the one-pel-shifted copy
        const int cw = cx >> 1;                                 // word offset of the column inside a window row
        uint32_t ps[8];
#pragma unroll
        for( int c = 0; c < 8; c++ ) ps[c] = 0u;
#pragma unroll 1
        for( int m = 0; m < 4; m++ )
        {
          const int bx8 = 2 * qx + ( m & 1 ), by8 = 2 * qy + ( m >> 1 ), b0 = 4 * q + m;
          const int sumA = sSumA[b0];
          uint32_t o[8][4];
#pragma unroll
          for( int y = 0; y < 8; y++ )
          {
            const uint4 ov = *reinterpret_cast<const uint4*>( org32 + ( by8 * 8 + y ) * ( R / 2 ) + bx8 * 4 );
            o[y][0] = ov.x; o[y][1] = ov.y; o[y][2] = ov.z; o[y][3] = ov.w;
          }
          int acc[8];
#pragma unroll
          for( int c = 0; c < 8; c++ ) acc[c] = sumA;
          const uint32_t* wp = wsrc + ( by8 * 8 + cy0 ) * wsw + bx8 * 4 + cw;
#pragma unroll
          for( int r = 0; r < 15; r++ )
          {
            uint32_t w[4];
            if( ( cw & 3 ) == 0 ) { const uint4 wv = *reinterpret_cast<const uint4*>( wp + r * wsw ); w[0] = wv.x; w[1] = wv.y; w[2] = wv.z; w[3] = wv.w; }
            else { w[0] = wp[r * wsw]; w[1] = wp[r * wsw + 1]; w[2] = wp[r * wsw + 2]; w[3] = wp[r * wsw + 3]; }
#pragma unroll
            for( int c = 0; c < 8; c++ )
            {
              if( r - c >= 0 && r - c < 8 )
              {
#pragma unroll
                for( int i = 0; i < 4; i++ ) acc[c] = __dp2a_lo( (int) __vmins2( o[r - c][i], w[i] ), (int) 0x0000fefeu, acc[c] );
              }
            }
          }
          const unsigned char* bb = bitsS + b0 * L.bStride;
          const uint32_t bx4 = 4u * bb[cx];
          const uint2 byw = *reinterpret_cast<const uint2*>( bb + nxp + cy0 );
          const uint16_t* vcol = V + ( by8 * 8 + cy0 ) * L.vPitch + bx8 * 8 + cx;
          uint32_t bk = 0xffffffffu;
#pragma unroll
          for( int c = 0; c < 8; c++ )
          {
            const uint32_t idx4 = __dp4a( c < 4 ? byw.x : byw.y, 1u << ( 8 * ( c & 3 ) ), bx4 );              // row bits are stored times 4
            const uint32_t mvk  = *reinterpret_cast<const uint32_t*>( sMv8 + c * ( PYR_MVN * 4 ) + idx4 );
            const uint32_t sad  = (uint32_t)( (int) vcol[c * L.vPitch] + acc[c] );
            ps[c] += sad;
            const uint32_t key = cy0 + c < ny ? sad * eight + mvk : 0xffffffffu;
            bk = min( bk, key );
          }
          uint32_t key = ( ( bk >> 3 ) << ob ) + (uint32_t)( ( cy0 + (int)( bk & 7u ) ) * nx + cx );
          if( uni ) { key = __reduce_min_sync( mask, key ); if( lane == lead ) atomicMin( &sKey32[b0], key ); }
          else atomicMin( &sKey32[b0], key );
        }
        {
          const unsigned char* bb = bitsS + ( OFF1 + q ) * L.bStride;
          const uint32_t bx4 = 4u * bb[cx];
          const uint2 byw = *reinterpret_cast<const uint2*>( bb + nxp + cy0 );
          uint32_t* tcol = LV >= 3 ? T + ( LV == 4 ? ( q >> 2 ) : 0 ) * L.tStride + cy0 * nxp + ( cx & 7 ) * nStrips + ( cx >> 3 ) : nullptr;
          uint32_t bk = 0xffffffffu;
#pragma unroll
          for( int c = 0; c < 8; c++ )
          {
            const uint32_t idx4 = __dp4a( c < 4 ? byw.x : byw.y, 1u << ( 8 * ( c & 3 ) ), bx4 );
            const uint32_t mvk  = *reinterpret_cast<const uint32_t*>( sMv8 + c * ( PYR_MVN * 4 ) + idx4 );
            if( cy0 + c < ny )
            {
              bk = min( bk, ps[c] * eight + mvk );
              if( LV >= 3 ) atomicAdd( tcol + c * nxp, ps[c] );
            }
          }
          uint32_t key = ( ( bk >> 3 ) << ob ) + (uint32_t)( ( cy0 + (int)( bk & 7u ) ) * nx + cx );
          if( uni ) { key = __reduce_min_sync( mask, key ); if( lane == lead ) atomicMin( &sKey32[OFF1 + q], key ); }
          else atomicMin( &sKey32[OFF1 + q], key );
        }
      }
    }
  }
  __syncthreads();

  // ---- 32x32 blocks from their tables, 64x64 root from the sum of the four tables
  if( LV >= 3 )
  {
    const int nTop = L.nT + ( LV == 4 ? 1 : 0 );
    const float invNx = 1.0f / (float) nx;
    for( int j = 0; j < nTop; j++ )
    {
      const bool isRoot64 = LV == 4 && j == L.nT;
      const int bid = isRoot64 ? OFF3 : OFF2 + j;
      const unsigned char* bb = bitsS + bid * L.bStride;
      unsigned long long best = ~0ull;
      for( int o = tid; o < nx * ny; o += nthr )
      {
        const int cy = fast_div( o, invNx ), cx = o - cy * nx;
        const int ti = cy * nxp + ( cx & 7 ) * nStrips + ( cx >> 3 );
        uint32_t s;
        if( isRoot64 ) s = T[ti] + T[L.tStride + ti] + T[2 * L.tStride + ti] + T[3 * L.tStride + ti];
        else           s = T[j * L.tStride + ti];
        const uint32_t bits = (uint32_t) bb[cx] + ( (uint32_t) bb[nxp + cy] >> 2 );
        const unsigned long long key = ( ( (unsigned long long) s + sMvRaw[bits < VVB_MVCOST_ENTRIES ? bits : VVB_MVCOST_ENTRIES - 1] ) << 16 ) | (unsigned) o;
        best = key < best ? key : best;
      }
#pragma unroll
      for( int mm = 16; mm > 0; mm >>= 1 ) { const unsigned long long o2 = __shfl_xor_sync( 0xffffffffu, best, mm ); best = o2 < best ? o2 : best; }
      if( lane == 0 && best != ~0ull ) atomicMin( &sKey64[j], best );
    }
    __syncthreads();
  }

  // ---- results
  for( int t = tid; t < NBLK; t += nthr )
  {
    const int l = t < OFF1 ? 0 : ( t < OFF2 ? 1 : ( t < OFF3 ? 2 : 3 ) );
    const int i = t - ( l == 0 ? 0 : ( l == 1 ? OFF1 : ( l == 2 ? OFF2 : OFF3 ) ) );
    unsigned long long cost; uint32_t order;
    if( l < 2 ) { const uint32_t k = sKey32[t]; cost = k >> ob; order = k & ( ( 1u << ob ) - 1u ); }
    else        { const unsigned long long k = sKey64[l == 2 ? i : L.nT]; cost = k >> 16; order = (uint32_t)( k & 0xffffu ); }
    const int cy = order / nx, cx = order - cy * nx;
    const unsigned char* bb = bitsS + t * L.bStride;
    const uint32_t bits = (uint32_t) bb[cx] + ( (uint32_t) bb[nxp + cy] >> 2 );
    vvb_best b;
    b.dx = (int16_t)( rb.left + cx ); b.dy = (int16_t)( rb.top + cy ); b.cost = cost;
    b.sad = (uint32_t)( cost - sMvRaw[bits < VVB_MVCOST_ENTRIES ? bits : VVB_MVCOST_ENTRIES - 1] );
    lv.best[l][( (size_t) root << ( 2 * ( LTOP - l ) ) ) + i] = b;
  }
}

} // namespace vvb
