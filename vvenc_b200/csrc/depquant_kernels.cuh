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


// ---------------------------------------------------------------------------------------------------------------------------------------------------------------
// dep_quant_quad_kernel: four lanes per TU, lane k owns trellis state k (and decision slot k).  Per scan position every lane prices its own state's two
// transitions (the "stay" candidate -- level A or zero -- and the "switch" candidate -- level B), the four decision slots gather their two candidates from the
// lanes the reference's call order names (checkRdCosts( 0, .., 0, 2 ), ( 1, .., 2, 0 ), ( 2, .., 1, 3 ), ( 3, .., 3, 1 ), DepQuant.cpp:1364-1367) with quad
// shuffles, and every lane then rebuilds its state from the lane its decision points to.  The 16-entry template / sum / level byte arrays of a state live in
// shared memory (word w of lane l at [w][l]: conflict-free, byte-addressable), everything else in registers.  The eight TUs of a warp walk the scan positions
// in lock step (a TU joins when the common position reaches its own first position), so the group-boundary work (update1StateEOS, CommonCtx::update) never
// diverges inside the warp and the ScanInfo record is a broadcast load.  Results equal dq_quant_tu (depquant_core.h) bit for bit.
#define VVB_DQQ_THREADS 128

__device__ __forceinline__ int dqq_byte( const uint32_t* a, int lane, int p ) { return ( a[( p >> 2 ) * VVB_DQQ_THREADS + lane] >> ( ( p & 3 ) * 8 ) ) & 255; }
__device__ __forceinline__ void dqq_set_byte( uint32_t* a, int lane, int p, int v )
{
  uint32_t& w = a[( p >> 2 ) * VVB_DQQ_THREADS + lane];
  const int sh = ( p & 3 ) * 8;
  w = ( w & ~( 255u << sh ) ) | ( (uint32_t)( v & 255 ) << sh );
}
__device__ __forceinline__ long long dqq_shfl64( long long v, int src )
{
  const int lo = __shfl_sync( 0xffffffffu, (int)( v & 0xffffffffll ), src ), hi = __shfl_sync( 0xffffffffu, (int)( v >> 32 ), src );
  return ( (long long) hi << 32 ) | (unsigned int) lo;
}

__global__ void __launch_bounds__( VVB_DQQ_THREADS ) dep_quant_quad_kernel( const __grid_constant__ DqLaunch L, const __grid_constant__ vvbdq::DqRates rates,
                                                                            const int32_t* __restrict__ coef, const uint8_t* __restrict__ needRdoq, int n,
                                                                            int16_t* __restrict__ q, int32_t* __restrict__ absSumOut, int32_t* __restrict__ lastPosOut, uint8_t* __restrict__ arena )
{
  using namespace vvbdq;
  __shared__ DqRates sR;
  __shared__ uint32_t sTpl[4 * VVB_DQQ_THREADS], sSum[4 * VVB_DQQ_THREADS], sAbs[4 * VVB_DQQ_THREADS];
  {
    const int32_t* src = reinterpret_cast<const int32_t*>( &rates );
    int32_t* dst = reinterpret_cast<int32_t*>( &sR );
    for( int i = threadIdx.x; i < (int)( sizeof( DqRates ) / 4 ); i += blockDim.x ) dst[i] = src[i];
  }
  __syncthreads();
  const int tid = threadIdx.x, lane = tid & 31, k = tid & 3, quadBase = lane & ~3;
  const unsigned quadMask = 0xfu << quadBase;
  const int slot = ( blockIdx.x * VVB_DQQ_THREADS + tid ) >> 2, nSlots = ( gridDim.x * VVB_DQQ_THREADS ) >> 2;
  const DqShape& sh = L.shape;
  const DqQuant& Q = L.quant;
  const int W = sh.width, H = sh.height, area = W * H, chunk = sh.numSbb + sh.numCoeff;
  uint8_t* ctxMem = arena + (size_t) slot * L.slotBytes;
  DqTrellis* trellis = reinterpret_cast<DqTrellis*>( ctxMem + L.ctxBytes );
  // geometry shared by all TUs of the launch (:1152-1172)
  bool zeroOut = false; int effW = W, effH = H;
  if( L.zeroOutMts ) { effH = H == 32 ? 16 : H; effW = W == 32 ? 16 : W; zeroOut = effH < H || effW < W; }
  const bool zeroOutForThres = zeroOut || 32 < H || 32 < W;
  const int zeroOutW = ( W == 32 && zeroOut ) ? 16 : 32, zeroOutH = ( H == 32 && zeroOut ) ? 16 : 32;
  int firstStart = min( W, 32 ) * min( H, 32 ) - 1;
  if( L.lfnst ) firstStart = ( ( W == 4 && H == 4 ) || ( W == 8 && H == 8 ) ) ? 7 : 15;
  const int defaultTh = Q.thresLast / (int)( Q.qScale << 2 );
  const int initRemRegBins = ( min( 32, effW ) * min( 32, effH ) * 28 ) / 16;

  for( int tu0 = 0; tu0 < n; tu0 += nSlots )                       // all lanes of the warp iterate together
  {
    const int tu = tu0 + slot;
    const bool have = tu < n;
    const bool skipTu = have && needRdoq && !needRdoq[tu];
    const int32_t* tc = coef + (size_t)( have ? tu : 0 ) * area;
    int16_t* qt = q + (size_t)( have ? tu : 0 ) * area;
    if( have ) for( int i = k; i < area; i += 4 ) qt[i] = 0;
    // ---- first position (findFirstPos, :58-73): the quad tests four positions per step
    int firstPos = -1;
    if( have && !skipTu )
    {
      for( int p0 = firstStart; p0 >= 0 && firstPos < 0; p0 -= 4 )
      {
        const int p = p0 - k;
        bool hit = false;
        if( p >= 0 )
        {
          const DqScanInfo& si = sh.scanInfo[p];
          if( !( zeroOutForThres && ( si.posX >= zeroOutW || si.posY >= zeroOutH ) ) ) hit = abs( tc[si.rasterPos] ) > defaultTh;
        }
        const unsigned b = __ballot_sync( quadMask, hit ) >> quadBase;
        if( b & 15u ) firstPos = p0 - ( __ffs( b & 15u ) - 1 );
      }
    }
    // ---- state of lane k (initStates, :682-695)
    long long rdCost = VVB_DQ_RDCOST_INIT, skipCost = VVB_DQ_RDCOST_INIT, dec0Cost = 0;
    int remRegBins = 4, skipRem = 4, sbb0 = 0, sbb1 = 0, skipSbb0 = 0, ctxSig = 0, ctxCff = 0, numSig = 0, refSbb = -1, ricePar = 0, riceZero = 0, currSet = 0;
    bool anyLt4 = true;
#pragma unroll
    for( int w = 0; w < 4; w++ ) { sTpl[w * VVB_DQQ_THREADS + tid] = 0; sSum[w * VVB_DQQ_THREADS + tid] = 0; sAbs[w * VVB_DQQ_THREADS + tid] = 0; }
    int warpFirst = firstPos;
#pragma unroll
    for( int m = 16; m >= 4; m >>= 1 ) warpFirst = max( warpFirst, __shfl_xor_sync( 0xffffffffu, warpFirst, m ) );

    for( int scanIdx = warpFirst; scanIdx >= 0; scanIdx-- )
    {
      const bool active = scanIdx <= firstPos;
      const DqScanInfo si = sh.scanInfo[scanIdx];
      const int spt = si.spt;
      const bool zo = zeroOut && ( si.posX >= effW || si.posY >= effH );
      // ---- xDecide (:1266-1386)
      long long decCost = VVB_DQ_RDCOST_INIT >> 2; int decLev = -1, decPrev = -2;
      if( !zo )
      {
        const int c = active ? tc[si.rasterPos] : 0;
        const long long scaledOrg = (long long) abs( c ) * Q.qScale;
        int qIdx = (int)( ( scaledOrg + Q.qAdd ) >> Q.qShift );
        const bool odd = qIdx < 0;
        long long distP, distQ, startDist; int levP, levQ, startLev; bool ge4 = false, rrg;
        if( odd )
        {
          const long long scaledAdd = Q.distStepAdd - scaledOrg * Q.distOrgFact;
          const long long distA = ( ( scaledAdd + 0 * Q.distStepAdd ) * 1 + Q.distAdd ) >> Q.distShift;
          const long long distB = ( ( scaledAdd + 1 * Q.distStepAdd ) * 2 + Q.distAdd ) >> Q.distShift;
          distQ = k < 2 ? distB : distA; levQ = 1; distP = 0; levP = 0;            // P: the zero transition, Q: level 1
          startDist = distB; startLev = 1;
          rrg = !anyLt4;
          if( anyLt4 && remRegBins < 4 )                                            // setRiceParam( k, scanInfo, prev, false ), :890-905
          {
            const int sumAll = max( min( 31, dqq_byte( sSum, tid, si.insidePos ) ), 0 );
            ricePar = c_goRicePars[sumAll]; riceZero = ( k < 2 ? 1 : 2 ) << ricePar;
          }
        }
        else
        {
          qIdx = max( 1, min( Q.maxQIdx, qIdx ) );
          const long long scaledAdd = qIdx * Q.distStepAdd - scaledOrg * Q.distOrgFact;
          // pqData[s]: j = ( s - qIdx ) & 3, level ( qIdx + j + 1 ) >> 1
          const int jP = ( ( k < 2 ? 0 : 3 ) - qIdx ) & 3, jQ = ( ( k < 2 ? 2 : 1 ) - qIdx ) & 3, j0 = ( 0 - qIdx ) & 3, j3 = ( 3 - qIdx ) & 3;
          distP = ( ( scaledAdd + jP * Q.distStepAdd ) * ( qIdx + jP ) + Q.distAdd ) >> Q.distShift; levP = ( qIdx + jP + 1 ) >> 1;
          distQ = ( ( scaledAdd + jQ * Q.distStepAdd ) * ( qIdx + jQ ) + Q.distAdd ) >> Q.distShift; levQ = ( qIdx + jQ + 1 ) >> 1;
          const bool cff02ge4 = ( ( qIdx + j0 + 1 ) >> 1 ) >= 4, cff13ge4 = ( ( qIdx + j3 + 1 ) >> 1 ) >= 4;
          ge4 = k < 2 ? cff02ge4 : cff13ge4;
          rrg = !( cff02ge4 || cff13ge4 || anyLt4 );
          if( ( anyLt4 || ge4 ) && ( remRegBins < 4 || ge4 ) )
          {
            const int sumAbs = dqq_byte( sSum, tid, si.insidePos );
            const int sumAll = max( min( 31, sumAbs - ( remRegBins < 4 ? 0 : 20 ) ), 0 );
            ricePar = c_goRicePars[sumAll];
            if( remRegBins < 4 ) riceZero = ( k < 2 ? 1 : 2 ) << ricePar;
          }
          // the start candidates use pqData[0] (slot 0) and pqData[2] (slot 2)
          const int jS = ( ( k == 0 ? 0 : 2 ) - qIdx ) & 3;
          startDist = ( ( scaledAdd + jS * Q.distStepAdd ) * ( qIdx + jS ) + Q.distAdd ) >> Q.distShift; startLev = ( qIdx + jS + 1 ) >> 1;
        }
        // ---- the two transitions out of state k (checkRdCosts :697-775 / checkRdCostsOdd1 :785-838)
        const int32_t* goRiceTab = c_goRiceBits[ricePar];
        long long costP, costQ, costZ = rdCost; int levPz = levP;
        costQ = rdCost + distQ; costP = rdCost + distP;
        if( rrg || remRegBins >= 4 )
        {
          const int32_t* cffBits = sR.gtxBits[ctxCff];
          const int32_t* sigBits = sR.sigBits[max( k - 1, 0 )][ctxSig];
          if( odd ) costQ += cffBits[1];
          else
          {
            if( levP < 4 ) costP += cffBits[levP]; else { const unsigned v = (unsigned)( levP - 4 ) >> 1; costP += cffBits[levP - ( v << 1 )] + goRiceTab[v < RICEMAX - 1 ? v : RICEMAX - 1]; }
            if( levQ < 4 ) costQ += cffBits[levQ]; else { const unsigned v = (unsigned)( levQ - 4 ) >> 1; costQ += cffBits[levQ - ( v << 1 )] + goRiceTab[v < RICEMAX - 1 ? v : RICEMAX - 1]; }
          }
          if( spt == SCAN_ISCSBB )      { costP += sigBits[1]; costQ += sigBits[1]; costZ += sigBits[0]; }
          else if( spt == SCAN_SOCSBB ) { costP += sbb1 + sigBits[1]; costQ += sbb1 + sigBits[1]; costZ += sbb1 + sigBits[0]; }
          else if( numSig )             { costP += sigBits[1]; costQ += sigBits[1]; costZ += sigBits[0]; }
          else costZ = VVB_DQ_RDCOST_INIT;
        }
        else
        {
          if( odd ) costQ += ( 1 << SCALE_BITS ) + goRiceTab[0];
          else
          {
            costP += ( 1 << SCALE_BITS ) + goRiceTab[levP <= riceZero ? levP - 1 : min( levP, RICEMAX - 1 )];
            costQ += ( 1 << SCALE_BITS ) + goRiceTab[levQ <= riceZero ? levQ - 1 : min( levQ, RICEMAX - 1 )];
          }
          costZ += goRiceTab[riceZero];
        }
        // "stay" candidate: level A if it beats zero, else zero (:756-767); in the odd branch it is the zero transition alone (:832-837)
        if( odd || !( costP < costZ ) ) { costP = costZ; levPz = 0; }
        // ---- decision slot k: first candidate from lane src1, second from lane src1 + 1 (the reference's call order); slots 0, 1 take "stay" first, slots 2, 3 "switch" first
        const int src1 = quadBase + ( ( k & 1 ) << 1 ), src2 = src1 + 1;
        const long long p1 = dqq_shfl64( costP, src1 ), q1 = dqq_shfl64( costQ, src1 ), p2 = dqq_shfl64( costP, src2 ), q2 = dqq_shfl64( costQ, src2 );
        const int lp1 = __shfl_sync( 0xffffffffu, levPz, src1 ), lq1 = __shfl_sync( 0xffffffffu, levQ, src1 ), lp2 = __shfl_sync( 0xffffffffu, levPz, src2 ), lq2 = __shfl_sync( 0xffffffffu, levQ, src2 );
        const long long f1 = k < 2 ? p1 : q1, f2 = k < 2 ? q2 : p2;
        const int l1 = k < 2 ? lp1 : lq1, l2 = k < 2 ? lq2 : lp2;
        if( f1 < decCost ) { decCost = f1; decLev = l1; decPrev = src1 - quadBase; }
        if( f2 < decCost ) { decCost = f2; decLev = l2; decPrev = src2 - quadBase; }
        // ---- checkRdCostStart (:848-869): slots 0 and 2 (odd branch: slot 2 only)
        if( ( k == 2 || ( k == 0 && !odd ) ) )
        {
          const int32_t* cffBits = sR.gtxBits[0];
          long long cst = startDist + ( sR.lastBitsX[si.posX] + sR.lastBitsY[si.posY] );
          if( startLev < 4 ) cst += cffBits[startLev];
          else { const unsigned v = (unsigned)( startLev - 4 ) >> 1; cst += cffBits[startLev - ( v << 1 )] + c_goRiceBits[0][v < RICEMAX ? v : RICEMAX - 1]; }
          if( cst < decCost ) { decCost = cst; decLev = startLev; decPrev = -1; }
        }
        if( spt == SCAN_EOCSBB )                                                       // checkRdCostSkipSbb (:871-880)
        {
          const long long cs = skipCost + skipSbb0;
          if( cs < decCost ) { decCost = cs; decLev = 0; decPrev = 4 | k; }
        }
      }
      else if( spt == SCAN_EOCSBB ) { decCost = skipCost + skipSbb0; decLev = 0; decPrev = 4 | k; }      // checkRdCostSkipSbbZeroOut (:882-888)

      if( active )
      {
        DqTrellis& t0 = trellis[2 * scanIdx];
        t0.absLevel[k] = (int16_t) decLev; t0.prevId[k] = (int8_t) decPrev;
        if( scanIdx && si.insidePos == 0 ) { DqTrellis& t1 = trellis[2 * scanIdx + 1]; t1.absLevel[k] = (int16_t) decLev; t1.prevId[k] = (int8_t) decPrev; }
        if( scanIdx == 0 ) dec0Cost = decCost;
      }
      if( scanIdx == 0 ) break;
      // ---- xDecideAndUpdate (:1396-1413): the shuffles below are executed by every lane, the results are kept by active TUs only
      if( active && spt == SCAN_SOCSBB ) { skipCost = rdCost; skipRem = remRegBins; skipSbb0 = sbb0; }
      const bool eos = si.insidePos == 0;
      if( eos || !zo )
      {
        const int srcLane = quadBase + ( decPrev >= 0 ? ( decPrev & 3 ) : k );
        const int pNumSig = __shfl_sync( 0xffffffffu, numSig, srcLane ), pRefSbb = __shfl_sync( 0xffffffffu, refSbb, srcLane ), pRem = __shfl_sync( 0xffffffffu, remRegBins, srcLane );
        const int pSbb0 = __shfl_sync( 0xffffffffu, sbb0, srcLane ), pSbb1 = __shfl_sync( 0xffffffffu, sbb1, srcLane ), pSkipRem = __shfl_sync( 0xffffffffu, skipRem, srcLane );
        const int srcTid = ( tid & ~31 ) + srcLane;
        uint32_t pT[4], pS[4], pA[4];
#pragma unroll
        for( int w = 0; w < 4; w++ ) { pT[w] = sTpl[w * VVB_DQQ_THREADS + srcTid]; pS[w] = sSum[w * VVB_DQQ_THREADS + srcTid]; pA[w] = sAbs[w * VVB_DQQ_THREADS + srcTid]; }
        __syncwarp();
        const bool upd = active && decPrev > -2;
        if( active ) { currSet ^= eos ? 4 : 0; rdCost = decCost; }
        if( upd )
        {
          const int lev = decLev, sub = lev < 2 ? lev : 3;
          if( !eos )                                                                   // update1State (:907-1000)
          {
            if( decPrev >= 0 )
            {
              numSig = pNumSig + ( lev ? 1 : 0 ); refSbb = pRefSbb; sbb0 = pSbb0; sbb1 = pSbb1;
              remRegBins = pRem - 1; if( remRegBins >= 4 ) remRegBins -= sub;
#pragma unroll
              for( int w = 0; w < 4; w++ ) { sTpl[w * VVB_DQQ_THREADS + tid] = pT[w]; sSum[w * VVB_DQQ_THREADS + tid] = pS[w]; sAbs[w * VVB_DQQ_THREADS + tid] = pA[w]; }
            }
            else
            {
              numSig = 1; refSbb = -1; remRegBins = initRemRegBins - sub;
#pragma unroll
              for( int w = 0; w < 4; w++ ) { sTpl[w * VVB_DQQ_THREADS + tid] = 0; sSum[w * VVB_DQQ_THREADS + tid] = 0; sAbs[w * VVB_DQQ_THREADS + tid] = 0; }
            }
            if( lev )
            {
              dqq_set_byte( sAbs, tid, si.insidePos, min( 126 + ( lev & 1 ), lev ) );
              const int min4or5 = min( 4 + ( lev & 1 ), lev );
              const int add = L.capSum ? min( 126 + ( lev & 1 ), lev ) : ( lev & 255 );
              for( int j = 0; j < si.numInv && j < 5; j++ )
              {
                const int p = si.invInPos[j];
                dqq_set_byte( sTpl, tid, p, dqq_byte( sTpl, tid, p ) + 32 + min4or5 );
                dqq_set_byte( sSum, tid, p, min( 255, dqq_byte( sSum, tid, p ) + add ) );
              }
            }
          }
          else                                                                         // update1StateEOS (:1002-1084)
          {
            uint32_t a[4];
            if( decPrev >= 4 )      { numSig = 0; remRegBins = pSkipRem; refSbb = decPrev - 4; a[0] = a[1] = a[2] = a[3] = 0; }
            else if( decPrev >= 0 ) { numSig = pNumSig + ( lev ? 1 : 0 ); refSbb = pRefSbb; remRegBins = pRem - 1; if( remRegBins >= 4 ) remRegBins -= sub; a[0] = pA[0]; a[1] = pA[1]; a[2] = pA[2]; a[3] = pA[3]; }
            else                    { numSig = 1; refSbb = -1; remRegBins = initRemRegBins - sub; a[0] = a[1] = a[2] = a[3] = 0; }
            a[0] = ( a[0] & ~255u ) | (uint32_t) min( 126 + ( lev & 1 ), lev );       // absVal[insidePos = 0]
            // levels of the finished group go to this state's buffer of the (already swapped) current set; the state's group arrays restart from zero
            uint8_t* flags  = ctxMem + (size_t)( currSet + k ) * chunk;
            uint8_t* levels = flags + sh.numSbb;
#pragma unroll
            for( int w = 0; w < 4; w++ )
#pragma unroll
              for( int b = 0; b < 4; b++ ) levels[scanIdx + 4 * w + b] = (uint8_t)( a[w] >> ( 8 * b ) );
            uint32_t tp[4] = { 0, 0, 0, 0 }, sm[4] = { 0, 0, 0, 0 };
            // CommonCtx::update (:473-531)
            const int maxDist = sh.nbOut[scanIdx - 1].maxDist;
            const int setCp = maxDist > 16 ? maxDist - 16 : 0;
            if( refSbb >= 0 )
            {
              const uint8_t* pf = ctxMem + (size_t)( ( currSet ^ 4 ) + refSbb ) * chunk; const uint8_t* pl = pf + sh.numSbb;
              for( int i = 0; i < sh.numSbb; i++ ) flags[i] = pf[i];
              for( int i = 0; i < setCp; i++ ) levels[scanIdx + 16 + i] = pl[scanIdx + 16 + i];
            }
            else
            {
              for( int i = 0; i < sh.numSbb; i++ ) flags[i] = 0;
              for( int i = 0; i < setCp; i++ ) levels[scanIdx + 16 + i] = 0;
            }
            flags[si.sbbPos] = numSig ? 1 : 0;
            const int sigNSbb = ( ( si.nextSbbRight ? flags[si.nextSbbRight] : 0 ) || ( si.nextSbbBelow ? flags[si.nextSbbBelow] : 0 ) ) ? 1 : 0;
            refSbb = k;
            sbb0 = sR.sigSbbBits[sigNSbb][0]; sbb1 = sR.sigSbbBits[sigNSbb][1];
            if( sigNSbb || ( ( si.nextSbbRight && si.nextSbbBelow ) ? flags[si.nextSbbBelow + 1] : 0 ) )
            {
              const DqNbOut* nb = sh.nbOut + ( scanIdx - 16 );
              const uint8_t* absLevels = levels + ( scanIdx - 16 );
              for( int id = 0; id < 16; id++, nb++ )
              {
                const int num = nb->num;
                if( num )
                {
                  int sumAbs = 0, sumAbs1 = 0, sumNum = 0;
                  for( int j = 0; j < num && j < 5; j++ ) { const int t = absLevels[nb->outPos[j]]; sumAbs += t; sumAbs1 += min( 4 + ( t & 1 ), t ); sumNum += t ? 1 : 0; }
                  tp[id >> 2] |= (uint32_t)( ( ( sumNum << 5 ) | sumAbs1 ) & 255 ) << ( ( id & 3 ) * 8 );
                  sm[id >> 2] |= (uint32_t) min( 255, sumAbs ) << ( ( id & 3 ) * 8 );
                }
              }
            }
#pragma unroll
            for( int w = 0; w < 4; w++ ) { sTpl[w * VVB_DQQ_THREADS + tid] = tp[w]; sSum[w * VVB_DQQ_THREADS + tid] = sm[w]; sAbs[w * VVB_DQQ_THREADS + tid] = 0; }
            numSig = 0;
          }
          if( remRegBins >= 4 )                                                        // the context part both updates end with (:986-998, 1070-1082)
          {
            const int t = dqq_byte( sTpl, tid, si.nextInsidePos );
            const int sumAbs1 = t & 31, sumNum = t >> 5;
            ctxSig = si.sigCtxOffsetNext + min( ( sumAbs1 + 1 ) >> 1, 3 );
            ctxCff = si.gtxCtxOffsetNext + min( sumAbs1 - sumNum, 4 );
          }
        }
        const unsigned lt = __ballot_sync( 0xffffffffu, upd && remRegBins < 4 );
        if( active ) anyLt4 = ( lt & quadMask ) != 0;
        __syncwarp();                                                                  // the group buffers written above are read by the other lanes of the quad at the next group boundary
      }
    }
    // ---- best path (:1238-1249) and backward scan (:1251-1262) by lane 0 of the quad
    long long c1 = dqq_shfl64( dec0Cost, quadBase + 1 ), c2 = dqq_shfl64( dec0Cost, quadBase + 2 ), c3 = dqq_shfl64( dec0Cost, quadBase + 3 );
    __syncwarp();
    if( have && k == 0 )
    {
      int absSum = 0, last = -1;
      if( firstPos >= 0 )
      {
        int prevId = -1; long long minPathCost = 0;
        if( dec0Cost < minPathCost ) { prevId = 0; minPathCost = dec0Cost; }
        if( c1 < minPathCost ) { prevId = 1; minPathCost = c1; }
        if( c2 < minPathCost ) { prevId = 2; minPathCost = c2; }
        if( c3 < minPathCost ) { prevId = 3; minPathCost = c3; }
        int scanIdx = 0;
        for( ; prevId >= 0; scanIdx++ )
        {
          if( prevId >= 4 && ( scanIdx & 15 ) ) continue;
          const DqTrellis& t = trellis[2 * scanIdx + ( prevId >> 2 )];
          const int absLevel = t.absLevel[prevId & 3];
          const int blkpos = sh.scanInfo[scanIdx].rasterPos;
          qt[blkpos] = (int16_t)( tc[blkpos] < 0 ? -absLevel : absLevel );
          absSum += absLevel;
          prevId = t.prevId[prevId & 3];
        }
        last = scanIdx - 1;
      }
      if( absSumOut ) absSumOut[tu] = absSum;
      if( lastPosOut ) lastPosOut[tu] = last;
    }
    __syncwarp();
  }
}

} // namespace vvb
