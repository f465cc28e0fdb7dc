// mctf_affine_kernels.cuh -- MCTF block-matching errors and affine-ME gradient helpers for sm_100a.
//
// MCTF: MCTF::motionErrorLuma (CommonLib/MCTF.cpp:1099-1164) -> motionErrorLumaInt (:122-145),
//       motionErrorLumaFrac6 (:147-203), motionErrorLumaFrac4 (:205-257); filter tables :72-110.
// Affine: xHorizontalSobelFilter / xVerticalSobelFilter / xEqualCoeffComputer (CommonLib/AffineGradientSearch.cpp:84-190).
#pragma once
#include "common.cuh"

namespace vvb {

// MCTF interpolation filters (constants of the algorithm, CommonLib/MCTF.cpp:72-110): row = 1/16-pel phase
__device__ __constant__ short c_mctfF8[16][8] = {
  {0,0,0,64,0,0,0,0},{0,1,-3,64,4,-2,0,0},{0,1,-6,62,9,-3,1,0},{0,2,-8,60,14,-5,1,0},{0,2,-9,57,19,-7,2,0},{0,3,-10,53,24,-8,2,0},
  {0,3,-11,50,29,-9,2,0},{0,3,-11,44,35,-10,3,0},{0,1,-7,38,38,-7,1,0},{0,3,-10,35,44,-11,3,0},{0,2,-9,29,50,-11,3,0},{0,2,-8,24,53,-10,3,0},
  {0,2,-7,19,57,-9,2,0},{0,1,-5,14,60,-8,2,0},{0,1,-3,9,62,-6,1,0},{0,0,-2,4,64,-3,1,0} };
__device__ __constant__ short c_mctfF4[16][4] = {
  {0,64,0,0},{-2,62,4,0},{-2,58,10,-2},{-4,56,14,-2},{-4,54,16,-2},{-6,52,20,-2},{-6,46,28,-4},{-4,42,30,-4},
  {-4,36,36,-4},{-4,30,42,-4},{-4,28,46,-6},{-2,20,52,-6},{-2,16,54,-4},{-2,14,56,-4},{-2,10,58,-2},{0,4,62,-2} };

#define MCTF_WARPS 4
#define MCTF_TMP_PITCH 64

// one warp per candidate; int32 error exactly as the reference (no early exit: besterror = INT_MAX, vvenc_unit_test.cpp:1552)
__global__ void __launch_bounds__( MCTF_WARPS * 32 ) mctf_error_kernel( const __grid_constant__ Plane orgPlane, const __grid_constant__ Plane refPlane,
                                                                        const vvb_mctf_cand* __restrict__ cands, int n, int tap4, int32_t* __restrict__ out )
{
  __shared__ short sTmp[MCTF_WARPS][( 64 + 6 ) * MCTF_TMP_PITCH];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warpsPerGrid = gridDim.x * MCTF_WARPS;
  const int maxv = ( 1 << refPlane.bitDepth ) - 1;
  for( int ci = blockIdx.x * MCTF_WARPS + warp; ci < n; ci += warpsPerGrid )
  {
    const vvb_mctf_cand c = cands[ci];
    const int w = c.w, h = c.h;
    int dx = c.mvx, dy = c.mvy;
    const int fx = dx & 15, fy = dy & 15;
    const int16_t* org = orgPlane.origin + (ptrdiff_t) c.y * orgPlane.stride + c.x;
    int err = 0;
    if( ( fx | fy ) == 0 )
    {
      dx /= 16; dy /= 16;                                  // MCTF.cpp:1121-1122 (C division, truncating)
      const int16_t* buf = refPlane.origin + (ptrdiff_t)( c.y + dy ) * refPlane.stride + c.x + dx;
      for( int i = lane; i < w * h; i += 32 )
      {
        const int y = i / w, x = i - y * w;
        const int d = (int) __ldg( org + (ptrdiff_t) y * orgPlane.stride + x ) - (int) __ldg( buf + (ptrdiff_t) y * refPlane.stride + x );
        err += d * d;
      }
    }
    else
    {
      dx >>= 4; dy >>= 4;                                  // MCTF.cpp:1136-1137 / :1151-1152 (arithmetic shift)
      const int16_t* buf = refPlane.origin + (ptrdiff_t)( c.y + dy ) * refPlane.stride + c.x + dx;
      const int taps = tap4 ? 4 : 6, first = tap4 ? 0 : 1, back = tap4 ? 1 : 2;
      const short* xf = tap4 ? c_mctfF4[fx] : c_mctfF8[fx];
      const short* yf = tap4 ? c_mctfF4[fy] : c_mctfF8[fy];
      short* tmp = sTmp[warp];
      const int rows = h + taps - 1;
      __syncwarp();
      for( int i = lane; i < rows * w; i += 32 )
      {
        const int r = i / w, x = i - r * w;
        const int16_t* p = buf + (ptrdiff_t)( r - back ) * refPlane.stride + x - back;
        int sum = 0;
        for( int t = 0; t < taps; t++ ) sum += xf[first + t] * (int) __ldg( p + t );
        sum = ( sum + 32 ) >> 6;
        tmp[r * MCTF_TMP_PITCH + x] = (short) min( max( sum, 0 ), maxv );
      }
      __syncwarp();
      for( int i = lane; i < w * h; i += 32 )
      {
        const int y = i / w, x = i - y * w;
        int sum = 0;
        for( int t = 0; t < taps; t++ ) sum += yf[first + t] * (int) tmp[( y + t ) * MCTF_TMP_PITCH + x];
        sum = ( sum + 32 ) >> 6;
        sum = min( max( sum, 0 ), maxv );
        const int d = sum - (int) __ldg( org + (ptrdiff_t) y * orgPlane.stride + x );
        err += d * d;
      }
    }
#pragma unroll
    for( int m = 16; m > 0; m >>= 1 ) err += __shfl_xor_sync( 0xffffffffu, err, m );
    if( lane == 0 ) out[ci] = err;
  }
}

// ---- affine: Sobel on a w x h prediction block with border replication (AffineGradientSearch.cpp:84-147)
__global__ void sobel_kernel( const int16_t* __restrict__ pred, int ps, int16_t* __restrict__ deriv, int ds, int w, int h, int vertical )
{
  for( int i = blockIdx.x * blockDim.x + threadIdx.x; i < w * h; i += gridDim.x * blockDim.x )
  {
    const int y = i / w, x = i - y * w;
    // border samples copy the nearest interior result (edges: inner neighbour, corners: inner diagonal)
    const int yy = min( max( y, 1 ), h - 2 ), xx = min( max( x, 1 ), w - 2 );
    const int16_t* c = pred + yy * ps + xx;
    int v;
    if( !vertical ) v = c[1 - ps] - c[-1 - ps] + ( c[1] << 1 ) - ( c[-1] << 1 ) + c[1 + ps] - c[-1 + ps];
    else            v = c[ps - 1] - c[-ps - 1] + ( c[ps] << 1 ) - ( c[-ps] << 1 ) + c[ps + 1] - c[-ps + 1];
    deriv[y * ds + x] = (int16_t) v;
  }
}

// ---- affine: normal-equation accumulation into int64[7][7] (AffineGradientSearch.cpp:150-190)
template<int NP>
__global__ void __launch_bounds__( 256 ) equal_coeff_kernel( const int16_t* __restrict__ resi, int rs, const int16_t* __restrict__ gx, const int16_t* __restrict__ gy, int ds,
                                                             int w, int h, long long* __restrict__ eq )
{
  long long acc[NP][NP + 1];
#pragma unroll
  for( int a = 0; a < NP; a++ )
#pragma unroll
    for( int b = 0; b <= NP; b++ ) acc[a][b] = 0;
  for( int i = blockIdx.x * blockDim.x + threadIdx.x; i < w * h; i += gridDim.x * blockDim.x )
  {
    const int j = i / w, k = i - j * w;
    const int cy = ( ( j >> 2 ) << 2 ) + 2, cx = ( ( k >> 2 ) << 2 ) + 2;
    const int a = gx[j * ds + k], b = gy[j * ds + k], r = resi[j * rs + k];
    int c[NP];
    if( NP == 4 ) { c[0] = a; c[1] = cx * a + cy * b; c[2] = b; c[3] = cy * a - cx * b; }
    else          { c[0] = a; c[1] = cx * a; c[2] = b; c[3] = cx * b; c[NP > 4 ? 4 : 0] = cy * a; c[NP > 4 ? 5 : 0] = cy * b; }
#pragma unroll
    for( int col = 0; col < NP; col++ )
    {
#pragma unroll
      for( int row = 0; row < NP; row++ ) acc[col][row] += (long long) c[col] * c[row];
      acc[col][NP] += ( (long long) c[col] * r ) * 8;
    }
  }
  __shared__ unsigned long long sEq[NP * ( NP + 1 )];
  for( int i = threadIdx.x; i < NP * ( NP + 1 ); i += blockDim.x ) sEq[i] = 0;
  __syncthreads();
#pragma unroll
  for( int col = 0; col < NP; col++ )
#pragma unroll
    for( int row = 0; row <= NP; row++ )
    {
      long long v = acc[col][row];
#pragma unroll
      for( int m = 16; m > 0; m >>= 1 ) v += __shfl_xor_sync( 0xffffffffu, v, m );
      if( ( threadIdx.x & 31 ) == 0 && v ) atomicAdd( &sEq[col * ( NP + 1 ) + row], (unsigned long long) v );
    }
  __syncthreads();
  for( int i = threadIdx.x; i < NP * ( NP + 1 ); i += blockDim.x )
  {
    const int col = i / ( NP + 1 ), row = i - col * ( NP + 1 );
    if( sEq[i] ) atomicAdd( reinterpret_cast<unsigned long long*>( &eq[( col + 1 ) * 7 + row] ), sEq[i] );
  }
}

} // namespace vvb
