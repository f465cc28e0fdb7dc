// trquant_kernels.cuh -- forward 2-D integer transform (DCT-II / DST-VII / DCT-VIII) + plain quantiser, fused.
//
// Replaces TrQuant::xT (CommonLib/TrQuant.cpp:481-564; 1-D cores CommonLib/TrQuant_EMT.cpp:366-421,1973-2000,
// AVX2 CommonLib/x86/TrafoX86.h:310-640) followed by Quant::quant -> QuantCore (CommonLib/Quant.cpp:735-833,132-230)
// and Quant::xNeedRDOQ -> needRdoqCore (:835-891,264-278) for luma TUs without LFNST / transform skip / scaling lists.
//
// Exactness: all sums are int32 exactly as the scalar reference.  Products run on IDP.2A (two int16 x int8 MACs per
// instruction): stage 1 always (residuals are int16), stage 2 when every stage-1 output fits int16 (always true for
// real video -- it is also the domain in which the AVX2 path, which saturates at TrafoX86.h:364, equals the scalar one);
// otherwise stage 2 falls back to plain 32-bit IMAD so that the result still equals the scalar reference.
//
// One *team* of T threads (4..128) owns one TU; a 128-thread CTA runs 128/T teams in lock step.
#pragma once
#include "common.cuh"

namespace vvb {

struct TuPar
{
  int w, h, lw, lh;
  int trHor, trVer;
  int keepW, keepH;          // non-zeroed-out outputs (TrQuant.cpp:496-497)
  int s1, s2;                // shifts (TrQuant.cpp:544-545)
  int offH, offV;            // offsets of the two matrices inside the int8 table (row j, column k)
  int scale, qbits;          // g_quantScales entry, iQBits (Quant.cpp:767-769)
  long long add, addRdoq;    // (171|85) << (qbits-9) ; 171 << (qbitsRdoq-9)
  int scaleRdoq, qbitsRdoq;
  int useThres;              // thres / (scale << 2)   (Quant.cpp:173-180, thrVal = 8)
  int scanOff;               // offset (entries) of this shape's raster->scanpos table
  int regionW, regionH;      // min(32,w), min(32,h)
  int team;                  // threads per TU
  int dqScale, dqShift;      // g_invQuantScales entry, rightShift of Quant::dequant (may be <= 0)   (Quant.cpp:554-561,601)
  int dqInMax;               // input clipping bound (Quant.cpp:606-607)
  int s2Inv;                 // second inverse shift 20 - bitDepth (TrQuant.cpp:609); the first is 7
  int pelMax;                // (1 << bitDepth) - 1, reconstruction clipping (Buffer.cpp:719)
  int lKeepW, lKeepH, lRegW; // log2 of keepW, keepH, regionW (all powers of two: every index split is a shift)
  int q32;                   // qbits <= 30: (|c| * scale + add) fits 32 bit for |c| < 2^16 (always true for residuals inside the bit depth)
  unsigned add32;            // low 32 bits of add (valid when q32)
  unsigned rdoqThr;          // smallest |c| with ((|c| * scaleRdoq + addRdoq) >> qbitsRdoq) != 0  (needRdoqCore as one compare)
};

// shared-memory carve-up per team (all in 32-bit words)
struct TeamSmem { int resiWords, tmpWords, coefWords, total; };

__host__ __device__ inline TeamSmem team_smem( const TuPar& p )
{
  TeamSmem s;
  s.resiWords = ( p.w * p.h ) / 2;                       // int16 residual, later reused for the int16 levels
  s.tmpWords  = p.keepW * p.h;                           // int32 or packed int16 stage-1 output [keepW][h]
  s.coefWords = p.regionW * p.regionH;                   // int32 coefficients of the scanned region
  s.total     = s.resiWords + s.tmpWords + s.coefWords + 8;
  return s;
}

// Matrix staging: Mt[q][j] (32-bit word) = bytes T[j][4q..4q+3]; j fastest so that a thread's 4 consecutive j are one LDS.128
__device__ __forceinline__ void stage_matrix( uint32_t* dst, const int8_t* __restrict__ table, int off, int N, int keep, int tid, int nthr )
{
  const int Q = N >> 2;
  for( int i = tid; i < Q * keep; i += nthr )
  {
    const int q = i / keep, j = i - q * keep;
    dst[i] = *reinterpret_cast<const uint32_t*>( table + off + j * N + 4 * q );
  }
}

// Views into one team's shared memory.
struct TeamView
{
  uint32_t* resi;     // int16 residual [h][w] as words; holds the int16 levels after team_forward
  uint32_t* tmp;      // stage-1 output [keepW][h] (int32)
  int32_t*  coef;     // int32 coefficients of the scanned region [regionH][regionW]
  int*      red;      // [0] ovf, [1] lastNZ, [2] cgLo, [3] cgHi, [4] absSum, [5] lastQ+1, [6] rdoq
};

__device__ __forceinline__ TeamView team_view( const TuPar& par, uint32_t* teamBase, int team )
{
  const int resiWords = ( par.w * par.h ) >> 1, tmpWords = par.keepW * par.h, coefWords = par.regionW * par.regionH;
  TeamView v;
  v.resi = teamBase + team * ( resiWords + tmpWords + coefWords + 8 );
  v.tmp  = v.resi + resiWords;
  v.coef = reinterpret_cast<int32_t*>( v.tmp + tmpWords );
  v.red  = reinterpret_cast<int*>( v.coef + coefWords );
  return v;
}

// Forward transform + quantiser of one TU by one team.  `load( i )` returns residual word i (two int16, row-major compact).
// Contains __syncthreads(): every thread of the CTA must call it, `live` masks the work.  On return (all threads synchronised)
// v.resi holds the levels, v.coef the coefficients, v.red[4] absSum, v.red[5] lastQ+1, v.red[6] the RDOQ flag; returns the final scan pos.
// lanes of the calling thread's team inside its warp (teams of 4..16 threads are aligned lane groups; larger teams span whole warps)
__device__ __forceinline__ unsigned team_lane_mask( int T )
{
  const unsigned lane = threadIdx.x & 31u;
  return T >= 32 ? 0xffffffffu : ( ( ( 1u << T ) - 1u ) << ( lane & ~(unsigned)( T - 1 ) ) );
}

// Plain quantiser of one TU by its team (Quant.cpp:132-230 QuantCore, :735-833 wrapper; needRdoqCore :264-278).
// coef: int32 [regionH][regionW] in shared memory; qWords: the level block int16 [h][w] (as words) in shared memory; inv: raster -> scan position.
// Every thread works on quads of 4 raster-consecutive coefficients (LDS.128 + one LDG.128 of scan positions); reductions are redux.sync inside
// the warp plus one shared atomic per warp for multi-warp teams.  Contains __syncthreads(); ends synchronised with red[4] = absSum,
// red[5] = last non-zero level's scan position + 1, red[6] = RDOQ flag; returns the final scan position (Quant.cpp:182-208).
__device__ __forceinline__ int team_quantise( const TuPar& par, const int32_t* coef, uint32_t* qWords, int* red, const int32_t* __restrict__ inv,
                                              int tt, int T, bool live )
{
  const unsigned tmask = team_lane_mask( T );
  const bool multi = T > 32;
  const int nQuads = live ? ( par.regionW * par.regionH ) >> 2 : 0;
  const int4* c4 = reinterpret_cast<const int4*>( coef );
  const int4* s4 = reinterpret_cast<const int4*>( inv );
  // ---- pass 1: last non-zero scan position, coefficient groups holding a value above the threshold, RDOQ pre-check
  int lastNZ = 0; unsigned cgLo = 0, cgHi = 0, rd = 0;
  for( int qi = tt; qi < nQuads; qi += T )
  {
    const int4 c = c4[qi];
    if( c.x | c.y | c.z | c.w )
    {
      const int4 sp = __ldg( s4 + qi );
#define VVB_Q1( cv, sv ) if( cv ) { const int ac = abs( cv ); lastNZ = max( lastNZ, sv ); rd |= (unsigned) ac >= par.rdoqThr; \
        if( ac > par.useThres ) { const int cg = ( sv ) >> 4; if( cg < 32 ) cgLo |= 1u << cg; else cgHi |= 1u << ( cg - 32 ); } }
      VVB_Q1( c.x, sp.x ) VVB_Q1( c.y, sp.y ) VVB_Q1( c.z, sp.z ) VVB_Q1( c.w, sp.w )
#undef VVB_Q1
    }
  }
  lastNZ = __reduce_max_sync( tmask, lastNZ );
  cgLo   = __reduce_or_sync( tmask, cgLo );
  cgHi   = __reduce_or_sync( tmask, cgHi );
  rd     = __reduce_or_sync( tmask, rd );
  if( multi )
  {
    if( ( threadIdx.x & 31 ) == 0 )
    {
      if( lastNZ ) atomicMax( &red[1], lastNZ );
      if( cgLo ) atomicOr( reinterpret_cast<unsigned*>( &red[2] ), cgLo );
      if( cgHi ) atomicOr( reinterpret_cast<unsigned*>( &red[3] ), cgHi );
      if( rd ) atomicOr( &red[6], 1 );
    }
    __syncthreads();
    lastNZ = red[1]; cgLo = (unsigned) red[2]; cgHi = (unsigned) red[3];
  }
  else if( tt == 0 ) red[6] = (int) rd;
  // ---- final scan position after trailing-CG trimming (Quant.cpp:182-208)
  int pos = lastNZ;
  {
    const int initCg = pos >> 4;
    if( initCg >= 1 )
    {
      const unsigned long long mask = ( (unsigned long long) cgHi << 32 ) | cgLo;
      const unsigned long long m = mask & ( initCg >= 63 ? ~0ull : ( ( 1ull << ( initCg + 1 ) ) - 1ull ) ) & ~1ull;   // CGs 1..initCg
      if( m == 0 ) pos = 15;
      else { const int g = 63 - __clzll( (long long) m ); if( g != initCg ) pos = g * 16 + 15; }
    }
  }
  // ---- quantise (Quant.cpp:211-227): levels of the scanned region, zeros elsewhere
  const int w = par.w;
  if( live && ( w > par.regionW || par.h > par.regionH ) )
    for( int i = tt; i < ( w * par.h ) >> 1; i += T )
    {
      const int y = i >> ( par.lw - 1 ), x = ( i & ( ( w >> 1 ) - 1 ) ) << 1;
      if( x >= par.regionW || y >= par.regionH ) qWords[i] = 0u;
    }
  int sum = 0, lastQ = 0;                                   // lastQ holds scan position + 1
  for( int qi = tt; qi < nQuads; qi += T )
  {
    const int4 c = c4[qi];
    int v0 = 0, v1 = 0, v2 = 0, v3 = 0;
    if( c.x | c.y | c.z | c.w )
    {
      const int4 sp = __ldg( s4 + qi );
#define VVB_Q2( cv, sv, vv ) if( ( cv ) && ( sv ) <= pos ) { \
        const unsigned ac = (unsigned) abs( cv ); \
        const int mag = ( par.q32 && ac < 65536u ) ? (int)( ( ac * (unsigned) par.scale + par.add32 ) >> par.qbits ) \
                                                   : (int)( ( (long long) ac * par.scale + par.add ) >> par.qbits ); \
        sum += mag; vv = min( 32767, mag ); if( ( cv ) < 0 ) vv = max( -32768, -mag ); if( vv ) lastQ = max( lastQ, ( sv ) + 1 ); }
      VVB_Q2( c.x, sp.x, v0 ) VVB_Q2( c.y, sp.y, v1 ) VVB_Q2( c.z, sp.z, v2 ) VVB_Q2( c.w, sp.w, v3 )
#undef VVB_Q2
    }
    const int idx = qi << 2, y = idx >> par.lRegW, x = idx & ( par.regionW - 1 );
    uint2 o;
    o.x = ( (uint32_t) v0 & 0xffffu ) | ( (uint32_t) v1 << 16 );
    o.y = ( (uint32_t) v2 & 0xffffu ) | ( (uint32_t) v3 << 16 );
    *reinterpret_cast<uint2*>( qWords + ( ( y * w + x ) >> 1 ) ) = o;
  }
  sum   = __reduce_add_sync( tmask, sum );
  lastQ = __reduce_max_sync( tmask, lastQ );
  if( multi )
  {
    if( ( threadIdx.x & 31 ) == 0 ) { if( sum ) atomicAdd( &red[4], sum ); if( lastQ ) atomicMax( &red[5], lastQ ); }
  }
  else if( tt == 0 ) { red[4] = sum; red[5] = lastQ; }
  __syncthreads();
  return pos;
}

template<class LOAD>
__device__ __forceinline__ int team_forward( const TuPar& par, const uint32_t* MtH, const uint32_t* MtV, const TeamView& v, const int32_t* __restrict__ scanTab,
                                             int tt, int T, bool live, LOAD load )
{
  const int w = par.w, h = par.h;
  const int resiWords = ( w * h ) >> 1, coefWords = par.regionW * par.regionH;
  uint32_t* myResi = v.resi; uint32_t* myTmp = v.tmp; int32_t* myCoef = v.coef; int* myRed = v.red;
  __syncthreads();                                           // previous iteration's smem fully consumed; matrices visible
  for( int i = tt; i < 8; i += T ) myRed[i] = 0;
  if( live )
    for( int i = tt; i < resiWords; i += T ) myResi[i] = load( i );
  __syncthreads();
  // ---- stage 1: tmp[j][i] = ( sum_k resi[i][k] * Th[j][k] + r1 ) >> s1   for i < h, j < keepW
  {
    const int lJG = par.lKeepW - 2, items = h << lJG;
    const int r1 = par.s1 > 0 ? 1 << ( par.s1 - 1 ) : 0;
    const int Q = w >> 2;
    int ovf = 0;
    int32_t* t = reinterpret_cast<int32_t*>( myTmp );
    for( int it = tt; live && it < items; it += T )
    {
      const int i = it >> lJG, j0 = ( it & ( ( 1 << lJG ) - 1 ) ) << 2;
      int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
      const uint2* rrow = reinterpret_cast<const uint2*>( myResi + ( i << ( par.lw - 1 ) ) );
      const uint32_t* mcol = MtH + j0;
      for( int q = 0; q < Q; q++ )
      {
        const uint2 rv = rrow[q];
        const uint4 m = *reinterpret_cast<const uint4*>( mcol + ( q << par.lKeepW ) );
        a0 = __dp2a_lo( (int) rv.x, (int) m.x, a0 ); a0 = __dp2a_hi( (int) rv.y, (int) m.x, a0 );
        a1 = __dp2a_lo( (int) rv.x, (int) m.y, a1 ); a1 = __dp2a_hi( (int) rv.y, (int) m.y, a1 );
        a2 = __dp2a_lo( (int) rv.x, (int) m.z, a2 ); a2 = __dp2a_hi( (int) rv.y, (int) m.z, a2 );
        a3 = __dp2a_lo( (int) rv.x, (int) m.w, a3 ); a3 = __dp2a_hi( (int) rv.y, (int) m.w, a3 );
      }
      a0 = ( a0 + r1 ) >> par.s1; a1 = ( a1 + r1 ) >> par.s1; a2 = ( a2 + r1 ) >> par.s1; a3 = ( a3 + r1 ) >> par.s1;
      ovf |= ( a0 != (short) a0 ) | ( a1 != (short) a1 ) | ( a2 != (short) a2 ) | ( a3 != (short) a3 );
      int32_t* td = t + ( j0 << par.lh ) + i;
      td[0] = a0; td[h] = a1; td[2 * h] = a2; td[3 * h] = a3;
    }
    if( ovf ) atomicOr( &myRed[0], 1 );
  }
  __syncthreads();
  // ---- stage 2: coef[j][i] = ( sum_k tmp[i][k] * Tv[j][k] + r2 ) >> s2   for i < keepW, j < keepH
  {
    const int lJG = par.lKeepH - 2, items = par.keepW << lJG;
    const int r2 = 1 << ( par.s2 - 1 );
    const int Q = h >> 2;
    const bool wide = myRed[0] != 0;
    const int32_t* t32 = reinterpret_cast<const int32_t*>( myTmp );
    for( int it = tt; live && it < items; it += T )
    {
      const int i = it >> lJG, j0 = ( it & ( ( 1 << lJG ) - 1 ) ) << 2;
      int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
      const int4* trow = reinterpret_cast<const int4*>( t32 + ( i << par.lh ) );
      const uint32_t* mcol = MtV + j0;
      if( !wide )
      {
        for( int q = 0; q < Q; q++ )
        {
          const int4 tv = trow[q];
          const uint32_t p0 = __byte_perm( (uint32_t) tv.x, (uint32_t) tv.y, 0x5410 );
          const uint32_t p1 = __byte_perm( (uint32_t) tv.z, (uint32_t) tv.w, 0x5410 );
          const uint4 m = *reinterpret_cast<const uint4*>( mcol + ( q << par.lKeepH ) );
          a0 = __dp2a_lo( (int) p0, (int) m.x, a0 ); a0 = __dp2a_hi( (int) p1, (int) m.x, a0 );
          a1 = __dp2a_lo( (int) p0, (int) m.y, a1 ); a1 = __dp2a_hi( (int) p1, (int) m.y, a1 );
          a2 = __dp2a_lo( (int) p0, (int) m.z, a2 ); a2 = __dp2a_hi( (int) p1, (int) m.z, a2 );
          a3 = __dp2a_lo( (int) p0, (int) m.w, a3 ); a3 = __dp2a_hi( (int) p1, (int) m.w, a3 );
        }
      }
      else
      {
        for( int q = 0; q < Q; q++ )
        {
          const int4 tv = trow[q];
          const uint4 m = *reinterpret_cast<const uint4*>( mcol + ( q << par.lKeepH ) );
#define VVB_MAC4( acc, mw ) acc += tv.x * (int)(signed char)( (mw) & 0xff ) + tv.y * (int)(signed char)( ( (mw) >> 8 ) & 0xff ) + tv.z * (int)(signed char)( ( (mw) >> 16 ) & 0xff ) + tv.w * (int)(signed char)( (mw) >> 24 )
          VVB_MAC4( a0, m.x ); VVB_MAC4( a1, m.y ); VVB_MAC4( a2, m.z ); VVB_MAC4( a3, m.w );
#undef VVB_MAC4
        }
      }
      a0 = ( a0 + r2 ) >> par.s2; a1 = ( a1 + r2 ) >> par.s2; a2 = ( a2 + r2 ) >> par.s2; a3 = ( a3 + r2 ) >> par.s2;
      int32_t* cd = myCoef + ( j0 << par.lRegW ) + i;
      cd[0] = a0; cd[par.regionW] = a1; cd[2 * par.regionW] = a2; cd[3 * par.regionW] = a3;
    }
    // rows/columns of the scanned region that were zeroed out (MTS 32 -> 16)
    if( live && ( par.keepW < par.regionW || par.keepH < par.regionH ) )
      for( int i = tt; i < coefWords; i += T )
      {
        const int y = i >> par.lRegW, x = i & ( par.regionW - 1 );
        if( x >= par.keepW || y >= par.keepH ) myCoef[i] = 0;
      }
  }
  __syncthreads();
  return team_quantise( par, myCoef, myResi, myRed, scanTab + par.scanOff, tt, T, live );
}

// results of team_forward -> global memory (q compact [h][w]; optional coefficients and per-TU scalars)
__device__ __forceinline__ void team_forward_store( const TuPar& par, const TeamView& v, int pos, int tu, int tt, int T, bool live,
                                                    int32_t* __restrict__ coefOut, int16_t* __restrict__ qOut, int32_t* __restrict__ absSumOut,
                                                    int32_t* __restrict__ lastPosOut, uint8_t* __restrict__ needRdoqOut )
{
  const int w = par.w, h = par.h, resiWords = ( w * h ) >> 1;
  const uint32_t* myResi = v.resi; const int32_t* myCoef = v.coef; const int* myRed = v.red;
  if( live )
  {
    uint32_t* dst = reinterpret_cast<uint32_t*>( qOut + (size_t) tu * w * h );
    for( int i = tt; i < resiWords; i += T ) dst[i] = myResi[i];
    if( coefOut )
    {
      int32_t* cd = coefOut + (size_t) tu * w * h;
      for( int i = tt; i < w * h; i += T )
      {
        const int y = i >> par.lw, x = i & ( w - 1 );
        cd[i] = ( x < par.regionW && y < par.regionH ) ? myCoef[( y << par.lRegW ) + x] : 0;
      }
    }
    if( tt == 0 )
    {
      const int sum = myRed[4];
      if( absSumOut )   absSumOut[tu]   = sum;
      if( lastPosOut )  lastPosOut[tu]  = sum ? myRed[5] - 1 : pos;      // Quant.cpp:806-816, :830
      if( needRdoqOut ) needRdoqOut[tu] = (uint8_t) myRed[6];
    }
  }
}

__global__ void __launch_bounds__( 128 ) fwd_trquant_kernel( const __grid_constant__ TuPar par, const int8_t* __restrict__ trTable, const int32_t* __restrict__ scanTab,
                                                             const int16_t* __restrict__ resi, int n,
                                                             int32_t* __restrict__ coefOut, int16_t* __restrict__ qOut, int32_t* __restrict__ absSumOut,
                                                             int32_t* __restrict__ lastPosOut, uint8_t* __restrict__ needRdoqOut )
{
  extern __shared__ __align__( 16 ) uint32_t smem[];
  const int T = par.team, nTeams = blockDim.x / T;
  const int team = threadIdx.x / T, tt = threadIdx.x - team * T;
  const int w = par.w, h = par.h;

  // ---- matrices, shared by all teams of the CTA
  uint32_t* MtH = smem;                                        // [w/4][keepW]
  uint32_t* MtV = MtH + ( w >> 2 ) * par.keepW;                // [h/4][keepH]
  uint32_t* teamBase = MtV + ( h >> 2 ) * par.keepH;
  stage_matrix( MtH, trTable, par.offH, w, par.keepW, threadIdx.x, blockDim.x );
  stage_matrix( MtV, trTable, par.offV, h, par.keepH, threadIdx.x, blockDim.x );
  const TeamView v = team_view( par, teamBase, team );

  for( int base = blockIdx.x * nTeams; base < n; base += gridDim.x * nTeams )
  {
    const int tu = base + team;
    const bool live = tu < n;
    const uint32_t* src = reinterpret_cast<const uint32_t*>( resi + (size_t)( live ? tu : 0 ) * w * h );
    const int pos = team_forward( par, MtH, MtV, v, scanTab, tt, T, live, [&]( int i ) { return __ldg( src + i ); } );
    team_forward_store( par, v, pos, tu, tt, T, live, coefOut, qOut, absSumOut, lastPosOut, needRdoqOut );
  }
}

// residual = org(x,y) - pred(x + start_x, y + start_y), written compactly so that fwd_trquant_kernel can consume it
__global__ void residual_from_planes_kernel( const __grid_constant__ Plane orgPlane, const __grid_constant__ Plane predPlane,
                                             const vvb_block* __restrict__ blocks, int n, int w, int h, int16_t* __restrict__ resi )
{
  const int area = w * h;
  const long long total = (long long) n * area;
  for( long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long) gridDim.x * blockDim.x )
  {
    const int b = (int)( i / area ), r = (int)( i - (long long) b * area );
    const int y = r / w, x = r - y * w;
    const vvb_block blk = blocks[b];
    const int o = __ldg( orgPlane.origin + (ptrdiff_t)( blk.y + y ) * orgPlane.stride + blk.x + x );
    const int p = __ldg( predPlane.origin + (ptrdiff_t)( blk.y + blk.start_y + y ) * predPlane.stride + blk.x + blk.start_x + x );
    resi[i] = (int16_t)( o - p );
  }
}

} // namespace vvb
