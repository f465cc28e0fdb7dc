// itrquant_kernels.cuh -- inverse path of the TU loop and the fused TU round trip.
//
//   inv_trquant_kernel   : TrQuant::invTransformNxN (CommonLib/TrQuant.cpp:318-348) = Quant::dequant (CommonLib/Quant.cpp:520-609,
//                          DeQuantCore :232-262) + TrQuant::xIT (:567-660; _fastInverseMM CommonLib/TrQuant_EMT.cpp:64-194, the B2..B8
//                          butterflies :231-636 equal the matrix product; AVX2 fastInvCore/roundClip CommonLib/x86/TrafoX86.h).
//   tu_roundtrip_kernel  : the luma TU candidate body of IntraSearch::xIntraCodingTUBlock (EncoderLib/IntraSearch.cpp:1328-1429) and of
//                          InterSearch::xEstimateInterResidualQT (EncoderLib/InterSearch.cpp:3659-3714): residual = org - pred,
//                          transformNxN, (absSum > 0 ? invTransformNxN : zero residual), PelBuf::reconstruct (CommonLib/Buffer.cpp:719),
//                          SSE(org, reco) / SSE(orgResi, recResi) / SSE(0, orgResi) -- one kernel, the levels never leave shared memory.
//
// Exactness: dequantised coefficients and first-pass outputs are clipped to 16 bit by the reference itself (transformMinimum/Maximum,
// clipMinimum/Maximum = -2^15 .. 2^15-1), so both inverse passes run on IDP.2A (two int16 x int8 MACs) with int32 sums exactly as the
// scalar code; sums cannot overflow (64 * 32768 * 90 < 2^31).
#pragma once
#include "trquant_kernels.cuh"

namespace vvb {

// Inverse matrices: dst[q*N + j] (word) = bytes T[4q][j], T[4q+1][j], T[4q+2][j], T[4q+3][j]   for q < keep/4  (k runs over coefficients)
__device__ __forceinline__ void stage_matrix_inv( uint32_t* dst, const int8_t* __restrict__ table, int off, int N, int keep, int tid, int nthr )
{
  const int Q = keep >> 2;
  for( int i = tid; i < Q * N; i += nthr )
  {
    const int q = i / N, j = i - q * N;
    const int8_t* t = table + off + ( 4 * q ) * N + j;
    dst[i] = (uint32_t)(uint8_t) t[0] | ( (uint32_t)(uint8_t) t[N] << 8 ) | ( (uint32_t)(uint8_t) t[2 * N] << 16 ) | ( (uint32_t)(uint8_t) t[3 * N] << 24 );
  }
}

__device__ __forceinline__ int clip16( int v ) { return max( -32768, min( 32767, v ) ); }

// words needed by team_inverse: cT [keepW][keepH/2 + 2] + tT [h][keepW/2]
__host__ __device__ inline int inv_ct_pitch( const TuPar& p ) { return ( p.keepH >> 1 ) + 2; }
__host__ __device__ inline int inv_words( const TuPar& p ) { return p.keepW * inv_ct_pitch( p ) + p.h * ( p.keepW >> 1 ); }

// Dequantise + inverse-transform one TU by one team.  qS: int16 levels [h][w] in shared memory; cT / tT: scratch (inv_words()).
// out( y, x0, r0, r1, r2, r3 ) receives the residual of row y, columns x0..x0+3.  Contains __syncthreads(): all threads of the CTA call it;
// `active` masks the work (a team whose TU quantised to zero, or a tail team, only walks the barriers).
template<class OUT>
__device__ __forceinline__ void team_inverse( const TuPar& par, const uint32_t* MvI, const uint32_t* MhI, const int16_t* qS, uint32_t* cT, uint32_t* tT,
                                              int tt, int T, bool active, OUT out )
{
  const int w = par.w, h = par.h, keepW = par.keepW, keepH = par.keepH;
  const int pitchC = inv_ct_pitch( par );
  // ---- dequant (DeQuantCore, Quant.cpp:232-262) + transpose: cT[i][k/2] = ( coef[k][i], coef[k+1][i] )
  if( active )
  {
    const int pairs = keepW * ( keepH >> 1 );
    const int sc = par.dqScale, sh = par.dqShift, inMax = par.dqInMax, inMin = -inMax - 1;
    for( int it = tt; it < pairs; it += T )
    {
      const int kp = it >> par.lKeepW, i = it & ( keepW - 1 );
      int c0 = max( inMin, min( inMax, (int) qS[( 2 * kp ) * w + i] ) );
      int c1 = max( inMin, min( inMax, (int) qS[( 2 * kp + 1 ) * w + i] ) );
      if( sh > 0 ) { const int add = 1 << ( sh - 1 ); c0 = ( c0 * sc + add ) >> sh; c1 = ( c1 * sc + add ) >> sh; }
      else         { c0 = (int)( (unsigned)( c0 * sc ) << ( -sh ) ); c1 = (int)( (unsigned)( c1 * sc ) << ( -sh ) ); }
      c0 = clip16( c0 ); c1 = clip16( c1 );
      cT[i * pitchC + kp] = ( (uint32_t) c0 & 0xffffu ) | ( (uint32_t) c1 << 16 );
    }
  }
  __syncthreads();
  // ---- pass 1 (vertical, shift 7): tmp[i][j] = clip16( ( sum_{k<keepH} coef[k][i] * Tv[k][j] + 64 ) >> 7 ), two columns i per item;
  //      stored transposed and packed: tT[j][i/2] = ( tmp[i][j], tmp[i+1][j] )
  if( active )
  {
    const int lJG = par.lh - 2, items = ( keepW >> 1 ) << lJG, Q = keepH >> 2, pitchT = keepW >> 1;
    for( int it = tt; it < items; it += T )
    {
      const int ip = it >> lJG, j0 = ( it & ( ( 1 << lJG ) - 1 ) ) << 2;
      int a0 = 0, a1 = 0, a2 = 0, a3 = 0, b0 = 0, b1 = 0, b2 = 0, b3 = 0;
      const uint2* ca = reinterpret_cast<const uint2*>( cT + ( 2 * ip ) * pitchC );
      const uint2* cb = reinterpret_cast<const uint2*>( cT + ( 2 * ip + 1 ) * pitchC );
      for( int q = 0; q < Q; q++ )
      {
        const uint2 va = ca[q], vb = cb[q];
        const uint4 m = *reinterpret_cast<const uint4*>( MvI + ( q << par.lh ) + j0 );
        a0 = __dp2a_lo( (int) va.x, (int) m.x, a0 ); a0 = __dp2a_hi( (int) va.y, (int) m.x, a0 );
        a1 = __dp2a_lo( (int) va.x, (int) m.y, a1 ); a1 = __dp2a_hi( (int) va.y, (int) m.y, a1 );
        a2 = __dp2a_lo( (int) va.x, (int) m.z, a2 ); a2 = __dp2a_hi( (int) va.y, (int) m.z, a2 );
        a3 = __dp2a_lo( (int) va.x, (int) m.w, a3 ); a3 = __dp2a_hi( (int) va.y, (int) m.w, a3 );
        b0 = __dp2a_lo( (int) vb.x, (int) m.x, b0 ); b0 = __dp2a_hi( (int) vb.y, (int) m.x, b0 );
        b1 = __dp2a_lo( (int) vb.x, (int) m.y, b1 ); b1 = __dp2a_hi( (int) vb.y, (int) m.y, b1 );
        b2 = __dp2a_lo( (int) vb.x, (int) m.z, b2 ); b2 = __dp2a_hi( (int) vb.y, (int) m.z, b2 );
        b3 = __dp2a_lo( (int) vb.x, (int) m.w, b3 ); b3 = __dp2a_hi( (int) vb.y, (int) m.w, b3 );
      }
#define VVB_P1( a, b ) ( ( (uint32_t) clip16( ( (a) + 64 ) >> 7 ) & 0xffffu ) | ( (uint32_t) clip16( ( (b) + 64 ) >> 7 ) << 16 ) )
      tT[( j0 + 0 ) * pitchT + ip] = VVB_P1( a0, b0 );
      tT[( j0 + 1 ) * pitchT + ip] = VVB_P1( a1, b1 );
      tT[( j0 + 2 ) * pitchT + ip] = VVB_P1( a2, b2 );
      tT[( j0 + 3 ) * pitchT + ip] = VVB_P1( a3, b3 );
#undef VVB_P1
    }
  }
  __syncthreads();
  // ---- pass 2 (horizontal, shift 20 - bitDepth): resi[y][x] = clip16( ( sum_{k<keepW} tmp[k][y] * Th[k][x] + rnd ) >> s2 )
  if( active )
  {
    const int lXG = par.lw - 2, items = h << lXG, Q = keepW >> 2, pitchT = keepW >> 1;
    const int s2 = par.s2Inv, r2 = 1 << ( s2 - 1 );
    for( int it = tt; it < items; it += T )
    {
      const int y = it >> lXG, x0 = ( it & ( ( 1 << lXG ) - 1 ) ) << 2;
      int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
      const uint2* tr = reinterpret_cast<const uint2*>( tT + y * pitchT );
      for( int q = 0; q < Q; q++ )
      {
        const uint2 tv = tr[q];
        const uint4 m = *reinterpret_cast<const uint4*>( MhI + ( q << par.lw ) + x0 );
        a0 = __dp2a_lo( (int) tv.x, (int) m.x, a0 ); a0 = __dp2a_hi( (int) tv.y, (int) m.x, a0 );
        a1 = __dp2a_lo( (int) tv.x, (int) m.y, a1 ); a1 = __dp2a_hi( (int) tv.y, (int) m.y, a1 );
        a2 = __dp2a_lo( (int) tv.x, (int) m.z, a2 ); a2 = __dp2a_hi( (int) tv.y, (int) m.z, a2 );
        a3 = __dp2a_lo( (int) tv.x, (int) m.w, a3 ); a3 = __dp2a_hi( (int) tv.y, (int) m.w, a3 );
      }
      out( y, x0, clip16( ( a0 + r2 ) >> s2 ), clip16( ( a1 + r2 ) >> s2 ), clip16( ( a2 + r2 ) >> s2 ), clip16( ( a3 + r2 ) >> s2 ) );
    }
  }
}

// shared memory of the inverse-only kernel: MvI [keepH/4][h] + MhI [keepW/4][w] + per team ( q [h][w] int16 + inv_words )
static inline size_t inv_trquant_smem( const TuPar& p, int nTeams )
{
  return ( (size_t)( p.keepH >> 2 ) * p.h + (size_t)( p.keepW >> 2 ) * p.w + (size_t) nTeams * ( ( p.w * p.h ) / 2 + inv_words( p ) ) ) * 4;
}

__global__ void __launch_bounds__( 128 ) inv_trquant_kernel( const __grid_constant__ TuPar par, const int8_t* __restrict__ trTable,
                                                             const int16_t* __restrict__ q, int n, int16_t* __restrict__ resiOut )
{
  extern __shared__ __align__( 16 ) uint32_t smem[];
  const int T = par.team, nTeams = blockDim.x / T;
  const int team = threadIdx.x / T, tt = threadIdx.x - team * T;
  const int w = par.w, h = par.h;
  uint32_t* MvI = smem;
  uint32_t* MhI = MvI + ( par.keepH >> 2 ) * h;
  uint32_t* teamBase = MhI + ( par.keepW >> 2 ) * w;
  stage_matrix_inv( MvI, trTable, par.offV, h, par.keepH, threadIdx.x, blockDim.x );
  stage_matrix_inv( MhI, trTable, par.offH, w, par.keepW, threadIdx.x, blockDim.x );
  const int qWords = ( w * h ) >> 1;
  uint32_t* myQ = teamBase + team * ( qWords + inv_words( par ) );
  uint32_t* cT  = myQ + qWords;
  uint32_t* tT  = cT + par.keepW * inv_ct_pitch( par );

  for( int base = blockIdx.x * nTeams; base < n; base += gridDim.x * nTeams )
  {
    const int tu = base + team;
    const bool live = tu < n;
    __syncthreads();
    if( live )
    {
      const uint32_t* src = reinterpret_cast<const uint32_t*>( q + (size_t) tu * w * h );
      for( int i = tt; i < qWords; i += T ) myQ[i] = __ldg( src + i );
    }
    __syncthreads();
    int16_t* dst = resiOut + (size_t)( live ? tu : 0 ) * w * h;
    team_inverse( par, MvI, MhI, reinterpret_cast<const int16_t*>( myQ ), cT, tT, tt, T, live,
                  [&]( int y, int x0, int r0, int r1, int r2, int r3 )
                  {
                    uint2 o;
                    o.x = ( (uint32_t) r0 & 0xffffu ) | ( (uint32_t) r1 << 16 );
                    o.y = ( (uint32_t) r2 & 0xffffu ) | ( (uint32_t) r3 << 16 );
                    *reinterpret_cast<uint2*>( dst + ( y << par.lw ) + x0 ) = o;
                  } );
  }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Fused TU round trip.  org / pred are either compact candidate pools [n][h][w] (orgPlane.origin == nullptr in the POOL instantiation)
// or positions inside resident planes (vvb_block: x, y, start_x/start_y = displacement of the prediction).
struct TuResult { unsigned long long distReco, distResi, distZero; int absSum, lastPos; };     // == vvb_tu_result (32 bytes)

// smem: forward matrices + inverse matrices + per team ( forward view ; the inverse scratch aliases v.tmp / v.coef )
static inline size_t tu_roundtrip_smem( const TuPar& p, int nTeams )
{
  const TeamSmem ts = team_smem( p );
  return ( (size_t)( p.w >> 2 ) * p.keepW + (size_t)( p.h >> 2 ) * p.keepH + (size_t)( p.keepH >> 2 ) * p.h + (size_t)( p.keepW >> 2 ) * p.w
           + (size_t) nTeams * ( ts.total + 8 ) ) * 4;
}

template<bool PLANES>
__global__ void __launch_bounds__( 128 ) tu_roundtrip_kernel( const __grid_constant__ TuPar par, const int8_t* __restrict__ trTable, const int32_t* __restrict__ scanTab,
                                                              const __grid_constant__ Plane orgPlane, const __grid_constant__ Plane predPlane, const vvb_block* __restrict__ blocks,
                                                              const int16_t* __restrict__ orgPool, const int16_t* __restrict__ predPool, int n,
                                                              int16_t* __restrict__ qOut, int16_t* __restrict__ recoOut, TuResult* __restrict__ resOut, uint8_t* __restrict__ needRdoqOut )
{
  extern __shared__ __align__( 16 ) uint32_t smem[];
  const int T = par.team, nTeams = blockDim.x / T;
  const int team = threadIdx.x / T, tt = threadIdx.x - team * T;
  const int w = par.w, h = par.h;
  uint32_t* MtH = smem;
  uint32_t* MtV = MtH + ( w >> 2 ) * par.keepW;
  uint32_t* MvI = MtV + ( h >> 2 ) * par.keepH;
  uint32_t* MhI = MvI + ( par.keepH >> 2 ) * h;
  uint32_t* teamBase = MhI + ( par.keepW >> 2 ) * w;
  stage_matrix( MtH, trTable, par.offH, w, par.keepW, threadIdx.x, blockDim.x );
  stage_matrix( MtV, trTable, par.offV, h, par.keepH, threadIdx.x, blockDim.x );
  stage_matrix_inv( MvI, trTable, par.offV, h, par.keepH, threadIdx.x, blockDim.x );
  stage_matrix_inv( MhI, trTable, par.offH, w, par.keepW, threadIdx.x, blockDim.x );
  const TeamSmem ts = team_smem( par );
  // team_view() strides teams by ts.total words; the extra 8 words per team (distortion accumulators) sit behind all views
  const TeamView v = team_view( par, teamBase, team );
  unsigned long long* acc = reinterpret_cast<unsigned long long*>( teamBase + nTeams * ts.total ) + team * 4;     // [0] reco, [1] resi, [2] zero
  const int hw = w >> 1;

  for( int base = blockIdx.x * nTeams; base < n; base += gridDim.x * nTeams )
  {
    const int tu = base + team;
    const bool live = tu < n;
    const int16_t* oBase; const int16_t* pBase; int so, sp;
    if( PLANES )
    {
      const vvb_block blk = blocks[live ? tu : 0];
      oBase = orgPlane.origin + (ptrdiff_t) blk.y * orgPlane.stride + blk.x;                 so = orgPlane.stride;
      pBase = predPlane.origin + (ptrdiff_t)( blk.y + blk.start_y ) * predPlane.stride + blk.x + blk.start_x;  sp = predPlane.stride;
    }
    else
    {
      oBase = orgPool + (size_t)( live ? tu : 0 ) * w * h;   so = w;
      pBase = predPool + (size_t)( live ? tu : 0 ) * w * h;  sp = w;
    }
    const int pos = team_forward( par, MtH, MtV, v, scanTab, tt, T, live, [&]( int i )
    {
      const int y = i >> ( par.lw - 1 ), x = ( i & ( hw - 1 ) ) << 1;
      const int16_t* o = oBase + (ptrdiff_t) y * so + x; const int16_t* p = pBase + (ptrdiff_t) y * sp + x;
      if( PLANES )
      {
        const int d0 = (int) __ldg( o ) - (int) __ldg( p ), d1 = (int) __ldg( o + 1 ) - (int) __ldg( p + 1 );
        return ( (uint32_t) d0 & 0xffffu ) | ( (uint32_t) d1 << 16 );
      }
      return __vsub2( __ldg( reinterpret_cast<const uint32_t*>( o ) ), __ldg( reinterpret_cast<const uint32_t*>( p ) ) );   // compact pools: words are aligned
    } );
    const int absSum = v.red[4];
    if( live )
    {
      uint32_t* dst = reinterpret_cast<uint32_t*>( qOut + (size_t) tu * w * h );
      for( int i = tt; i < ( w * h ) >> 1; i += T ) dst[i] = v.resi[i];
      if( tt == 0 ) { acc[0] = 0; acc[1] = 0; acc[2] = 0; }
    }
    // every thread has read absSum before any thread can pass the first barrier of team_inverse; red[] is only reset in the next team_forward
    const bool active = live && absSum > 0;
    unsigned long long dReco = 0, dResi = 0, dZero = 0;
    const int pelMax = par.pelMax;
    int16_t* rBase = recoOut ? recoOut + (size_t)( live ? tu : 0 ) * w * h : nullptr;
    auto account = [&]( int y, int x0, int r0, int r1, int r2, int r3 )
    {
      const int16_t* o = oBase + (ptrdiff_t) y * so + x0; const int16_t* p = pBase + (ptrdiff_t) y * sp + x0;
      const int r[4] = { r0, r1, r2, r3 };
      int rc[4], ovs[4], pvs[4];
      if( !PLANES || ( ( ( reinterpret_cast<uintptr_t>( o ) | reinterpret_cast<uintptr_t>( p ) ) & 7 ) == 0 ) )
      {
        const uint2 ow = __ldg( reinterpret_cast<const uint2*>( o ) ), pw = __ldg( reinterpret_cast<const uint2*>( p ) );
        ovs[0] = lo16( ow.x ); ovs[1] = hi16( ow.x ); ovs[2] = lo16( ow.y ); ovs[3] = hi16( ow.y );
        pvs[0] = lo16( pw.x ); pvs[1] = hi16( pw.x ); pvs[2] = lo16( pw.y ); pvs[3] = hi16( pw.y );
      }
      else
      {
#pragma unroll
        for( int c = 0; c < 4; c++ ) { ovs[c] = __ldg( o + c ); pvs[c] = __ldg( p + c ); }
      }
#pragma unroll
      for( int c = 0; c < 4; c++ )
      {
        const int ov = ovs[c], pv = pvs[c];
        rc[c] = max( 0, min( pelMax, pv + r[c] ) );
        const int dz = ov - pv;                 // original residual
        const long long dr = (long long) dz - r[c];
        const int dc = ov - rc[c];
        dZero += (unsigned) ( dz * dz );
        dResi += (unsigned long long)( dr * dr );
        dReco += (unsigned) ( dc * dc );
      }
      if( rBase )
      {
        uint2 ov2;
        ov2.x = ( (uint32_t) rc[0] & 0xffffu ) | ( (uint32_t) rc[1] << 16 );
        ov2.y = ( (uint32_t) rc[2] & 0xffffu ) | ( (uint32_t) rc[3] << 16 );
        *reinterpret_cast<uint2*>( rBase + ( y << par.lw ) + x0 ) = ov2;
      }
    };
    // inverse scratch aliases the forward tmp / coef areas (both dead once the levels are in v.resi)
    team_inverse( par, MvI, MhI, reinterpret_cast<const int16_t*>( v.resi ), v.tmp, v.tmp + par.keepW * inv_ct_pitch( par ), tt, T, active, account );
    if( live && !active )                       // quantised to zero: residual 0 (IntraSearch.cpp:1366-1369 piResi.fill(0))
    {
      const int lXG = par.lw - 2, items = h << lXG;
      for( int it = tt; it < items; it += T ) account( it >> lXG, ( it & ( ( 1 << lXG ) - 1 ) ) << 2, 0, 0, 0, 0 );
    }
    // team reduction: shuffles inside the warp (teams of 4..16 lanes are aligned lane groups), then one shared atomic per warp and value
    {
      const int span = T < 32 ? T : 32;
      for( int off = span >> 1; off > 0; off >>= 1 )
      {
        dReco += __shfl_xor_sync( 0xffffffffu, dReco, off );
        dResi += __shfl_xor_sync( 0xffffffffu, dResi, off );
        dZero += __shfl_xor_sync( 0xffffffffu, dZero, off );
      }
      if( live && ( tt & ( span - 1 ) ) == 0 )
      {
        if( T <= 32 ) { acc[0] = dReco; acc[1] = dResi; acc[2] = dZero; }
        else { atomicAdd( &acc[0], dReco ); atomicAdd( &acc[1], dResi ); atomicAdd( &acc[2], dZero ); }
      }
    }
    __syncthreads();
    if( live && tt == 0 )
    {
      TuResult r;
      r.distReco = acc[0]; r.distResi = acc[1]; r.distZero = acc[2];
      r.absSum = absSum; r.lastPos = absSum ? v.red[5] - 1 : pos;
      resOut[tu] = r;
      if( needRdoqOut ) needRdoqOut[tu] = (uint8_t) v.red[6];
    }
  }
}

} // namespace vvb
