// depquant_kernels.cuh -- DepQuant::xQuantDQ on the device: one thread walks the trellis of one TU (depquant_core.h holds the algorithm and the reference
// line numbers).  The four trellis states of a TU depend on one another at every scan position, and the positions are strictly sequential, so the parallelism
// is across TUs: a picture's worth of TUs of one shape per launch.  Rate tables and scan geometry are staged in shared memory once per CTA; the per-thread
// working set (the eight level buffers of CommonCtx and the 12-byte trellis records) lives in a global arena indexed by thread slot, not by TU.
#pragma once
#include "common.cuh"
#include "depquant_core.h"

namespace vvb {

struct DqLaunch
{
  vvbdq::DqShape shape;              // device pointers
  vvbdq::DqQuant quant;
  int32_t zeroOutMts, lfnst, capSum;
  uint32_t ctxBytes, slotBytes;      // bytes of the CommonCtx memory (rounded to 16) and of one thread slot in the arena
};

#define VVB_DQ_THREADS 64

__global__ void __launch_bounds__( VVB_DQ_THREADS ) dep_quant_kernel( const __grid_constant__ DqLaunch L, const __grid_constant__ vvbdq::DqRates rates,
                                                                      const int32_t* __restrict__ coef, const uint8_t* __restrict__ needRdoq, int n,
                                                                      int16_t* __restrict__ q, int32_t* __restrict__ absSum, int32_t* __restrict__ lastPos, uint8_t* __restrict__ arena )
{
  __shared__ vvbdq::DqRates sRates;
  {
    const int32_t* src = reinterpret_cast<const int32_t*>( &rates );
    int32_t* dst = reinterpret_cast<int32_t*>( &sRates );
    for( int i = threadIdx.x; i < (int)( sizeof( vvbdq::DqRates ) / 4 ); i += blockDim.x ) dst[i] = src[i];
  }
  __syncthreads();
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  vvbdq::DqWork wk;
  wk.ctxMem  = arena + (size_t) slot * L.slotBytes;
  wk.trellis = reinterpret_cast<vvbdq::DqTrellis*>( wk.ctxMem + L.ctxBytes );
  const int area = L.shape.width * L.shape.height;
  for( int tu = slot; tu < n; tu += gridDim.x * blockDim.x )
  {
    int16_t* qt = q + (size_t) tu * area;
    int32_t sum = 0, last = -1;
    if( needRdoq && !needRdoq[tu] ) { for( int i = 0; i < area; i++ ) qt[i] = 0; }       // DepQuant::quant, :1464-1468 (useSelectiveRdoq)
    else vvbdq::dq_quant_tu( L.shape, L.quant, sRates, L.zeroOutMts != 0, L.lfnst != 0, L.capSum != 0, coef + (size_t) tu * area, qt, wk, &sum, &last );
    if( absSum ) absSum[tu] = sum;
    if( lastPos ) lastPos[tu] = last;
  }
}

} // namespace vvb
