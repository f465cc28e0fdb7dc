// rdoq_kernels.cuh -- QuantRDOQ2::xRateDistOptQuantFast on the device: one thread decides the levels of one TU (rdoq_core.h holds the algorithm and the reference
// line numbers).  Inside a TU the decisions are strictly sequential -- the context of a coefficient is a function of the levels chosen for its five already-visited
// neighbours, the budget of context-coded bins runs along the scan, the last-position optimisation looks at running sums -- so the parallelism is across TUs: a
// picture's worth of TUs of one shape per launch.  Rate tables and the scan order are staged in shared memory once per CTA; the level buffer a thread reads its
// templates from is its TU's slice of the output (global memory, L1-resident for the group being worked on); the five 16-entry cost arrays of the current
// coefficient group are thread-local.
#pragma once
#include "common.cuh"
#include "rdoq_core.h"

namespace vvb {

struct RqLaunch
{
  vvbrq::RqPar   par;
  const int32_t* scan;               // device: scan position -> raster index inside the scanned region (ctx->d_scan, second half), min(32,w) * min(32,h) entries
  int32_t        numScan;
};

#define VVB_RQ_THREADS 64

__global__ void __launch_bounds__( VVB_RQ_THREADS ) rdoq_kernel( const __grid_constant__ RqLaunch L, const __grid_constant__ vvbrq::RqRates rates,
                                                                 const int32_t* __restrict__ coef, const uint8_t* __restrict__ needRdoq, int n,
                                                                 int16_t* __restrict__ q, int32_t* __restrict__ absSum, int32_t* __restrict__ lastPos )
{
  __shared__ vvbrq::RqRates sRates;
  __shared__ int32_t sScan[1024];
  {
    const int32_t* src = reinterpret_cast<const int32_t*>( &rates );
    int32_t* dst = reinterpret_cast<int32_t*>( &sRates );
    for( int i = threadIdx.x; i < (int)( sizeof( vvbrq::RqRates ) / 4 ); i += blockDim.x ) dst[i] = src[i];
    for( int i = threadIdx.x; i < L.numScan; i += blockDim.x ) sScan[i] = L.scan[i];
  }
  __syncthreads();
  const int area = L.par.width * L.par.height;
  for( int tu = blockIdx.x * blockDim.x + threadIdx.x; tu < n; tu += gridDim.x * blockDim.x )
  {
    int16_t* qt = q + (size_t) tu * area;
    int32_t sum = 0, last = -1;
    if( needRdoq && !needRdoq[tu] ) { for( int i = 0; i < area; i++ ) qt[i] = 0; }       // QuantRDOQ2::quant, :273, 291-295 (useSelectiveRdoq)
    else vvbrq::rq_quant_tu( L.par, sRates, sScan, coef + (size_t) tu * area, qt, &sum, &last );
    if( absSum ) absSum[tu] = sum;
    if( lastPos ) lastPos[tu] = last;
  }
}

// second engine (rq_quant_tu_v2): the same launch shape; the cost tables arrive as a third by-value parameter, the group index map is derived from the staged scan order
__global__ void __launch_bounds__( VVB_RQ_THREADS ) rdoq_v2_kernel( const __grid_constant__ RqLaunch L, const __grid_constant__ vvbrq::RqRates rates, const __grid_constant__ vvbrq::RqCost cost,
                                                                    const int32_t* __restrict__ coef, const uint8_t* __restrict__ needRdoq, int n,
                                                                    int16_t* __restrict__ q, int32_t* __restrict__ absSum, int32_t* __restrict__ lastPos )
{
  __shared__ vvbrq::RqRates sRates;
  __shared__ vvbrq::RqCost sCost;
  __shared__ int32_t sScan[1024];
  __shared__ uint8_t sCgIdx[64];
  {
    const int32_t* src = reinterpret_cast<const int32_t*>( &rates );
    int32_t* dst = reinterpret_cast<int32_t*>( &sRates );
    for( int i = threadIdx.x; i < (int)( sizeof( vvbrq::RqRates ) / 4 ); i += blockDim.x ) dst[i] = src[i];
    const int32_t* srcC = reinterpret_cast<const int32_t*>( &cost );
    int32_t* dstC = reinterpret_cast<int32_t*>( &sCost );
    for( int i = threadIdx.x; i < (int)( sizeof( vvbrq::RqCost ) / 4 ); i += blockDim.x ) dstC[i] = srcC[i];
    for( int i = threadIdx.x; i < L.numScan; i += blockDim.x ) sScan[i] = L.scan[i];
    if( threadIdx.x < 64 ) sCgIdx[threadIdx.x] = 0;
  }
  __syncthreads();
  {
    const int wg = L.par.regionW >> 2, lrw = ( L.par.regionW == 32 ? 5 : L.par.regionW == 16 ? 4 : L.par.regionW == 8 ? 3 : 2 );
    for( int g = threadIdx.x; g < ( L.numScan >> 4 ); g += blockDim.x )
    {
      const int r = sScan[g << 4], x = r & ( L.par.regionW - 1 ), y = r >> lrw;
      sCgIdx[( y >> 2 ) * wg + ( x >> 2 )] = (uint8_t) g;
    }
  }
  __syncthreads();
  const int area = L.par.width * L.par.height;
  for( int tu = blockIdx.x * blockDim.x + threadIdx.x; tu < n; tu += gridDim.x * blockDim.x )
  {
    int16_t* qt = q + (size_t) tu * area;
    int32_t sum = 0, last = -1;
    if( needRdoq && !needRdoq[tu] ) { for( int i = 0; i < area; i++ ) qt[i] = 0; }
    else vvbrq::rq_quant_tu_v2( L.par, sRates, sCost, sScan, sCgIdx, coef + (size_t) tu * area, qt, &sum, &last );
    if( absSum ) absSum[tu] = sum;
    if( lastPos ) lastPos[tu] = last;
  }
}

// transform-skip variant (rq_ts_quant_tu): the same shape -- one thread per TU, rate tables and scan order in shared memory, the level buffer is the output slice
struct RqTsLaunch
{
  vvbrq::RqTsPar par;
  const int32_t* scan;
  int32_t        numScan;
};

__global__ void __launch_bounds__( VVB_RQ_THREADS ) rdoq_ts_kernel( const __grid_constant__ RqTsLaunch L, const __grid_constant__ vvbrq::RqTsRates rates,
                                                                    const int32_t* __restrict__ coef, const uint8_t* __restrict__ needRdoq, int n,
                                                                    int16_t* __restrict__ q, int32_t* __restrict__ absSum )
{
  __shared__ vvbrq::RqTsRates sRates;
  __shared__ int32_t sScan[1024];
  {
    const int32_t* src = reinterpret_cast<const int32_t*>( &rates );
    int32_t* dst = reinterpret_cast<int32_t*>( &sRates );
    for( int i = threadIdx.x; i < (int)( sizeof( vvbrq::RqTsRates ) / 4 ); i += blockDim.x ) dst[i] = src[i];
    for( int i = threadIdx.x; i < L.numScan; i += blockDim.x ) sScan[i] = L.scan[i];
  }
  __syncthreads();
  const int area = L.par.width * L.par.height;
  for( int tu = blockIdx.x * blockDim.x + threadIdx.x; tu < n; tu += gridDim.x * blockDim.x )
  {
    int16_t* qt = q + (size_t) tu * area;
    int32_t sum = 0;
    if( needRdoq && !needRdoq[tu] ) { for( int i = 0; i < area; i++ ) qt[i] = 0; }
    else vvbrq::rq_ts_quant_tu( L.par, sRates, sScan, coef + (size_t) tu * area, qt, &sum );
    if( absSum ) absSum[tu] = sum;
  }
}

// BDPCM variant (rq_bdpcm_quant_tu): the reconstruction the next position predicts from (w * h int32 per TU) lives in a global arena indexed by thread slot, like the
// level buffers of the DepQuant kernel
__global__ void __launch_bounds__( VVB_RQ_THREADS ) rdoq_bdpcm_kernel( const __grid_constant__ RqTsLaunch L, const __grid_constant__ vvbrq::RqBdpcmPar B, const __grid_constant__ vvbrq::RqTsRates rates,
                                                                       const int32_t* __restrict__ coef, const uint8_t* __restrict__ needRdoq, int n,
                                                                       int16_t* __restrict__ q, int32_t* __restrict__ absSum, int32_t* __restrict__ arena )
{
  __shared__ vvbrq::RqTsRates sRates;
  __shared__ int32_t sScan[1024];
  {
    const int32_t* src = reinterpret_cast<const int32_t*>( &rates );
    int32_t* dst = reinterpret_cast<int32_t*>( &sRates );
    for( int i = threadIdx.x; i < (int)( sizeof( vvbrq::RqTsRates ) / 4 ); i += blockDim.x ) dst[i] = src[i];
    for( int i = threadIdx.x; i < L.numScan; i += blockDim.x ) sScan[i] = L.scan[i];
  }
  __syncthreads();
  const int area = L.par.width * L.par.height;
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  int32_t* full = arena + (size_t) slot * area;
  for( int tu = slot; tu < n; tu += gridDim.x * blockDim.x )
  {
    int16_t* qt = q + (size_t) tu * area;
    int32_t sum = 0;
    if( needRdoq && !needRdoq[tu] ) { for( int i = 0; i < area; i++ ) qt[i] = 0; }
    else vvbrq::rq_bdpcm_quant_tu( L.par, B, sRates, sScan, coef + (size_t) tu * area, qt, full, &sum );
    if( absSum ) absSum[tu] = sum;
  }
}

} // namespace vvb
