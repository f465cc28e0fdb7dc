// trquant_tc_kernels.cuh -- forward 2-D integer transform on the 5th-generation tensor cores (tcgen05, kind::i8, TMEM
// accumulators) + the same fused quantiser as trquant_kernels.cuh.  Square TUs 16x16, 32x32, 64x64.
//
// Exact-integer strategy: the transform is pure int32 (TrQuant_EMT.cpp:1973-2000), the matrices fit s8
// (|T| <= 91), so both stages run as s8 x s8 -> s32 MMAs on byte planes of the left operand:
//   stage 1:  r   = r1*2^7  + r0               (r0 in [0,127], r1 = r >> 7, |r| < 2^14)        -> 2 MMAs per K step
//   stage 2:  tmp = t2*2^14 + t1*2^7 + t0      (t0,t1 in [0,127], t2 = tmp >> 14, |tmp| < 2^21) -> 3 MMAs per K step
// and the epilogue recombines  sum = (d2 << 14) + (d1 << 7) + d0  in int32 before the rounding shift.  No value is ever
// rounded, so the coefficients equal the scalar reference bit for bit (also where the AVX2 path would saturate).
//
// Tile = 128 stacked rows = 128/N TUs.  A operands are written to shared memory by the threads themselves in the canonical
// K-major no-swizzle layout ([16-byte K chunk][row][16 B]: SBO = 128 B, LBO = rows*16 B), B = the transform matrix rows in
// the same layout, D lives in TMEM (lane = stacked row, column = output index).  Stage-1 results are scattered transposed
// (bytes) into the stage-2 A operand, so the second transform is again "rows x matrix".
#pragma once
#include "common.cuh"
#include "trquant_kernels.cuh"

namespace vvb {

__device__ __forceinline__ uint32_t smem_u32( const void* p ) { return (uint32_t) __cvta_generic_to_shared( p ); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE (cute/arch/mma_sm100_desc.hpp SmemDescriptor; cute/atom/mma_traits_sm100.hpp:273-303)
__device__ __forceinline__ uint64_t umma_desc_kmajor( uint32_t smemAddr, uint32_t lboBytes, uint32_t sboBytes )
{
  uint64_t d = 0;
  d |= (uint64_t)( ( smemAddr >> 4 ) & 0x3fffu );           // start address, bits [0,14)
  d |= (uint64_t)( ( lboBytes >> 4 ) & 0x3fffu ) << 16;     // leading byte offset, bits [16,30)
  d |= (uint64_t)( ( sboBytes >> 4 ) & 0x3fffu ) << 32;     // stride byte offset, bits [32,46)
  d |= (uint64_t) 1 << 46;                                   // version = 1 (sm_100)
  return d;                                                  // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

// UMMA instruction descriptor for kind::i8, s8 x s8 -> s32, both operands K-major (mma_sm100_desc.hpp InstrDescriptor)
__device__ __forceinline__ uint32_t umma_idesc_i8( int M, int N )
{
  uint32_t d = 0;
  d |= 2u << 4;                      // c_format = S32
  d |= 1u << 7;                      // a_format = INT8 (signed)
  d |= 1u << 10;                     // b_format = INT8 (signed)
  d |= (uint32_t)( N >> 3 ) << 17;   // n_dim
  d |= (uint32_t)( M >> 4 ) << 24;   // m_dim
  return d;
}

__device__ __forceinline__ void umma_i8( uint32_t tmemD, uint64_t descA, uint64_t descB, uint32_t idesc, uint32_t accumulate )
{
  asm volatile(
    "{\n\t"
    ".reg .pred p;\n\t"
    "setp.ne.b32 p, %4, 0;\n\t"
    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
    "}\n"
    :: "r"( tmemD ), "l"( descA ), "l"( descB ), "r"( idesc ), "r"( accumulate ), "r"( 0u ), "r"( 0u ), "r"( 0u ), "r"( 0u ) : "memory" );
}

__device__ __forceinline__ void umma_commit( uint32_t mbarAddr )
{
  asm volatile( "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"( mbarAddr ) : "memory" );
}

__device__ __forceinline__ void mbar_init( uint32_t addr, uint32_t count ) { asm volatile( "mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"( addr ), "r"( count ) : "memory" ); }

__device__ __forceinline__ void mbar_wait( uint32_t addr, uint32_t parity )
{
  uint32_t done = 0;
  while( !done )
  {
    asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n" : "=r"( done ) : "r"( addr ), "r"( parity ) : "memory" );
  }
}

__device__ __forceinline__ void tmem_ld16( uint32_t taddr, int (&v)[16] )
{
  asm volatile( "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"( v[0] ), "=r"( v[1] ), "=r"( v[2] ), "=r"( v[3] ), "=r"( v[4] ), "=r"( v[5] ), "=r"( v[6] ), "=r"( v[7] ),
                  "=r"( v[8] ), "=r"( v[9] ), "=r"( v[10] ), "=r"( v[11] ), "=r"( v[12] ), "=r"( v[13] ), "=r"( v[14] ), "=r"( v[15] )
                : "r"( taddr ) : "memory" );
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile( "tcgen05.wait::ld.sync.aligned;" ::: "memory" ); }
__device__ __forceinline__ void tc_fence_before() { asm volatile( "tcgen05.fence::before_thread_sync;" ::: "memory" ); }
__device__ __forceinline__ void tc_fence_after()  { asm volatile( "tcgen05.fence::after_thread_sync;" ::: "memory" ); }
__device__ __forceinline__ void fence_async_smem() { asm volatile( "fence.proxy.async.shared::cta;" ::: "memory" ); }

#define TC_TMEM_COLS 128

// N = TU size (16, 32, 64).  KB = bytes of K per operand row = max(32, N); KEEP = N > 32 ? 32 : N (zero-out, DCT-II only at 64).
template<int N>
__global__ void __launch_bounds__( 128 ) fwd_trquant_tc_kernel( const __grid_constant__ TuPar par, const int8_t* __restrict__ trTable, const int32_t* __restrict__ scanTab,
                                                                const int16_t* __restrict__ resi, int n,
                                                                int32_t* __restrict__ coefOut, int16_t* __restrict__ qOut, int32_t* __restrict__ absSumOut,
                                                                int32_t* __restrict__ lastPosOut, uint8_t* __restrict__ needRdoqOut )
{
  constexpr int KB   = N < 32 ? 32 : N;          // operand row length in bytes (K elements, zero padded to a multiple of 32)
  constexpr int NCH  = KB / 16;                  // 16-byte K chunks
  constexpr int TPT  = 128 / N;                  // TUs per 128-row tile
  constexpr int KEEP = N > 32 ? 32 : N;          // kept outputs per dimension for DCT-II
  constexpr int A_BYTES = NCH * 128 * 16;        // one byte-plane operand of 128 rows
  constexpr int B_BYTES = NCH * 32 * 16 * ( KEEP > 32 ? 2 : 1 );   // matrix rows (<= 32 kept rows, padded to 32)
  constexpr int REGION = KEEP * KEEP;            // scanned coefficients per TU

  extern __shared__ __align__( 128 ) unsigned char smemTc[];
  unsigned char* sA   = smemTc;                              // 3 byte planes (stage 1 uses 2)
  unsigned char* sBh  = sA + 3 * A_BYTES;                    // horizontal matrix, rows j < keepW
  unsigned char* sBv  = sBh + B_BYTES;                       // vertical matrix, rows j < keepH
  int32_t*       sCoef = reinterpret_cast<int32_t*>( sBv + B_BYTES );          // [TPT][REGION]
  uint32_t*      sQ    = reinterpret_cast<uint32_t*>( sCoef + TPT * REGION );  // [TPT][N*N/2] int16 pairs (levels)
  int*           sRed  = reinterpret_cast<int*>( sQ + TPT * N * N / 2 );       // [TPT][8]
  __shared__ __align__( 8 ) unsigned long long sMbar;
  __shared__ uint32_t sTmemBase;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int keepW = par.keepW, keepH = par.keepH;            // == KEEP for DCT-II; 16 for DST-VII/DCT-VIII at 32
  const uint32_t mbar = smem_u32( &sMbar );

  // ---- one-time set-up: TMEM allocation (warp 0), barrier init, matrices in canonical layout
  if( warp == 0 )
  {
    asm volatile( "tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"( smem_u32( &sTmemBase ) ), "r"( (uint32_t) TC_TMEM_COLS ) : "memory" );
    asm volatile( "tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory" );
  }
  if( tid == 0 ) { mbar_init( mbar, 1 ); asm volatile( "fence.mbarrier_init.release.cluster;" ::: "memory" ); }
  // B[chunk c][row j (0..31)][16 B] = T[j][16c .. 16c+15], zero beyond N (K padding) and beyond the kept rows
  for( int i = tid; i < NCH * 32 * 16; i += 128 )
  {
    const int c = i / ( 32 * 16 ), r = ( i / 16 ) % 32, b = i % 16, k = c * 16 + b;
    sBh[i] = ( r < keepW && k < N ) ? (unsigned char) trTable[par.offH + r * N + k] : 0;
    sBv[i] = ( r < keepH && k < N ) ? (unsigned char) trTable[par.offV + r * N + k] : 0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sTmemBase;
  uint32_t phase = 0;

  const uint32_t idesc = umma_idesc_i8( 128, 32 );           // N operand padded to 32 columns for every TU size (rows >= keep are zero)
  const uint32_t aAddr = smem_u32( sA ), bhAddr = smem_u32( sBh ), bvAddr = smem_u32( sBv );
  const int numTiles = ( n + TPT - 1 ) / TPT;
  const int r1 = par.s1 > 0 ? 1 << ( par.s1 - 1 ) : 0, r2 = 1 << ( par.s2 - 1 );
  const int tuInTile = tid / N, rowInTu = tid % N;
  const int32_t* inv = scanTab + par.scanOff;

  for( int tile = blockIdx.x; tile < numTiles; tile += gridDim.x )
  {
    const int tu = tile * TPT + tuInTile;
    const bool live = tu < n;
    // ---- stage-1 A operand: thread = stacked row; byte planes r0 = r & 127, r1 = r >> 7
    {
      uint32_t w0[KB / 4], w1[KB / 4];
#pragma unroll
      for( int i = 0; i < KB / 4; i++ ) { w0[i] = 0; w1[i] = 0; }
      if( live )
      {
        const uint4* src = reinterpret_cast<const uint4*>( resi + ( (size_t) tu * N + rowInTu ) * N );
#pragma unroll
        for( int v = 0; v < N / 8; v++ )
        {
          const uint4 q = __ldg( src + v );
          const uint32_t ww[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
          for( int j = 0; j < 4; j++ )
          {
            const int e0 = lo16( ww[j] ), e1 = hi16( ww[j] );
            const int k = v * 8 + 2 * j;                     // element index of e0
            w0[k / 4] |= (uint32_t)( e0 & 127 ) << ( 8 * ( k & 3 ) );        w0[( k + 1 ) / 4] |= (uint32_t)( e1 & 127 ) << ( 8 * ( ( k + 1 ) & 3 ) );
            w1[k / 4] |= (uint32_t)( ( e0 >> 7 ) & 255 ) << ( 8 * ( k & 3 ) ); w1[( k + 1 ) / 4] |= (uint32_t)( ( e1 >> 7 ) & 255 ) << ( 8 * ( ( k + 1 ) & 3 ) );
          }
        }
      }
#pragma unroll
      for( int c = 0; c < NCH; c++ )
      {
        *reinterpret_cast<uint4*>( sA + 0 * A_BYTES + ( c * 128 + tid ) * 16 ) = make_uint4( w0[4*c], w0[4*c+1], w0[4*c+2], w0[4*c+3] );
        *reinterpret_cast<uint4*>( sA + 1 * A_BYTES + ( c * 128 + tid ) * 16 ) = make_uint4( w1[4*c], w1[4*c+1], w1[4*c+2], w1[4*c+3] );
      }
      for( int i = tid; i < TPT * 8; i += 128 ) sRed[i] = 0;
    }
    fence_async_smem();
    __syncthreads();
    // ---- stage-1 MMAs: D_p[128 x 32] (columns 32p..) = A_p[128 x KB] * Bh^T, p = 0,1
    if( tid == 0 )
    {
      tc_fence_after();
#pragma unroll
      for( int p = 0; p < 2; p++ )
#pragma unroll
        for( int ks = 0; ks < KB / 32; ks++ )
        {
          const uint64_t da = umma_desc_kmajor( aAddr + p * A_BYTES + ks * 2 * 128 * 16, 128 * 16, 128 );
          const uint64_t db = umma_desc_kmajor( bhAddr + ks * 2 * 32 * 16, 32 * 16, 128 );
          umma_i8( tmem + 32 * p, da, db, idesc, ks > 0 ? 1u : 0u );
        }
      umma_commit( mbar );
    }
    mbar_wait( mbar, phase ); phase ^= 1;
    tc_fence_after();
    // ---- stage-1 epilogue: tmp[i][j] = ((d1 << 7) + d0 + r1) >> s1 ; scatter three byte planes, transposed, as stage-2 A
    {
      const uint32_t lane = (uint32_t)( warp * 32 ) << 16;
      // zero the three planes first (rows of TUs that keep fewer columns, K padding) -- each thread clears its own rows
#pragma unroll
      for( int p = 0; p < 3; p++ )
#pragma unroll
        for( int c = 0; c < NCH; c++ ) *reinterpret_cast<uint4*>( sA + p * A_BYTES + ( c * 128 + tid ) * 16 ) = make_uint4( 0, 0, 0, 0 );
      int d0[16], d1[16];
      int tmpv[32];
#pragma unroll
      for( int half = 0; half < 2; half++ )
      {
        tmem_ld16( tmem + lane + 0  + 16 * half, d0 );
        tmem_ld16( tmem + lane + 32 + 16 * half, d1 );
        tmem_ld_wait();
#pragma unroll
        for( int j = 0; j < 16; j++ ) tmpv[16 * half + j] = ( ( d1[j] << 7 ) + d0[j] + r1 ) >> par.s1;
      }
      tc_fence_before();
      __syncthreads();                                   // every row's old A bytes are consumed and cleared before the scatter
      if( live )
      {
        // stage-2 stacked row = tuInTile * keepW + j ; K index = rowInTu
        const int c = rowInTu >> 4, b = rowInTu & 15;
#pragma unroll
        for( int j = 0; j < 32; j++ )
        {
          if( j < keepW )
          {
            const int t = tmpv[j];
            const int row2 = tuInTile * keepW + j;
            unsigned char* dst = sA + ( c * 128 + row2 ) * 16 + b;
            dst[0 * A_BYTES] = (unsigned char)( t & 127 );
            dst[1 * A_BYTES] = (unsigned char)( ( t >> 7 ) & 127 );
            dst[2 * A_BYTES] = (unsigned char)( ( t >> 14 ) & 255 );
          }
        }
      }
    }
    fence_async_smem();
    __syncthreads();
    // ---- stage-2 MMAs: D_p[128 x 32] = A2_p * Bv^T, p = 0,1,2
    if( tid == 0 )
    {
      tc_fence_after();
#pragma unroll
      for( int p = 0; p < 3; p++ )
#pragma unroll
        for( int ks = 0; ks < KB / 32; ks++ )
        {
          const uint64_t da = umma_desc_kmajor( aAddr + p * A_BYTES + ks * 2 * 128 * 16, 128 * 16, 128 );
          const uint64_t db = umma_desc_kmajor( bvAddr + ks * 2 * 32 * 16, 32 * 16, 128 );
          umma_i8( tmem + 32 * p, da, db, idesc, ks > 0 ? 1u : 0u );
        }
      umma_commit( mbar );
    }
    mbar_wait( mbar, phase ); phase ^= 1;
    tc_fence_after();
    // ---- stage-2 epilogue: stacked row = (TU t2, column i') ; coef[j'][i'] = ((d2<<14) + (d1<<7) + d0 + r2) >> s2
    {
      const uint32_t lane = (uint32_t)( warp * 32 ) << 16;
      const int t2 = tid / keepW, i2 = tid - t2 * keepW;
      const bool rowLive = t2 < TPT && ( tile * TPT + t2 ) < n;
      int d0[16], d1[16], d2[16];
#pragma unroll
      for( int half = 0; half < 2; half++ )
      {
        tmem_ld16( tmem + lane + 0  + 16 * half, d0 );
        tmem_ld16( tmem + lane + 32 + 16 * half, d1 );
        tmem_ld16( tmem + lane + 64 + 16 * half, d2 );
        tmem_ld_wait();
        if( rowLive )
        {
#pragma unroll
          for( int j = 0; j < 16; j++ )
          {
            const int jj = 16 * half + j;
            if( jj < keepH ) sCoef[t2 * REGION + jj * KEEP + i2] = ( ( d2[j] << 14 ) + ( d1[j] << 7 ) + d0[j] + r2 ) >> par.s2;
          }
        }
      }
      // MTS at 32 keeps 16x16 of the 32x32 scan region: clear the rest
      if( keepW < KEEP || keepH < KEEP )
        for( int i = tid; i < TPT * REGION; i += 128 )
        {
          const int rr = ( i % REGION ) / KEEP, cc = i % KEEP;
          if( cc >= keepW || rr >= keepH ) sCoef[i] = 0;
        }
    }
    tc_fence_before();
    __syncthreads();

    // ---- quantiser: the same device function as the CUDA-core kernel, team of N threads per TU
    {
      const int tt = rowInTu; constexpr int T = N;
      int32_t*  myCoef = sCoef + tuInTile * REGION;
      int*      myRed  = sRed + tuInTile * 8;
      uint32_t* myQ    = sQ + tuInTile * ( N * N / 2 );
      const int pos = team_quantise<( N == 16 ? 4 : N == 32 ? 5 : 6 ), ( N == 16 ? 4 : N == 32 ? 5 : 6 ), N>( par, myCoef, myQ, myRed, inv, tt, live );
      if( live )
      {
        uint32_t* dst = reinterpret_cast<uint32_t*>( qOut + (size_t) tu * N * N );
        for( int i = tt; i < N * N / 2; i += T ) dst[i] = myQ[i];
        if( coefOut )
        {
          int32_t* cd = coefOut + (size_t) tu * N * N;
          for( int i = tt; i < N * N; i += T )
          {
            const int y = i / N, x = i - y * N;
            cd[i] = ( x < KEEP && y < KEEP ) ? myCoef[y * KEEP + x] : 0;
          }
        }
        if( tt == 0 )
        {
          const int sum = myRed[4];
          if( absSumOut )   absSumOut[tu]   = sum;
          if( lastPosOut )  lastPosOut[tu]  = sum ? myRed[5] - 1 : pos;
          if( needRdoqOut ) needRdoqOut[tu] = (uint8_t) myRed[6];
        }
      }
    }
    __syncthreads();
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if( warp == 0 ) asm volatile( "tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"( tmem ), "r"( (uint32_t) TC_TMEM_COLS ) : "memory" );
}

template<int N> static inline size_t trquant_tc_smem()
{
  constexpr int KB = N < 32 ? 32 : N, NCH = KB / 16, TPT = 128 / N, KEEP = N > 32 ? 32 : N;
  return (size_t) 3 * NCH * 128 * 16 + 2 * (size_t) NCH * 32 * 16 + (size_t) TPT * KEEP * KEEP * 4 + (size_t) TPT * N * N * 2 + TPT * 8 * 4 + 256;
}

} // namespace vvb
