// depquant_core.h -- dependent (trellis-coded) quantisation of one luma TU: DepQuant::xQuantDQ (CommonLib/DepQuant.cpp:1129-1264) with its helpers
// xDecide / xDecideAndUpdate (:1266-1414), checkRdCosts / checkRdCostsOdd1 / checkRdCostStart / checkRdCostSkipSbb (:697-888), setRiceParam (:890-905),
// update1State / update1StateEOS (:907-1084) and CommonCtx::update (:473-531), restated for one thread per TU.
//
// What stays on the host: everything that depends on the encoder's entropy-coding state.  The rate tables RateEstimator::initCtx derives from the CABAC
// contexts (:344-471: last-position bits, significant-group bits, significance bits of the three context sets, greater-than bits) arrive as `DqRates`;
// the quantiser constants of Quantizer::initQuantBlock (:533-572, double arithmetic on lambda) arrive as `DqQuant`; the scan geometry of
// Rom::xInitScanArrays / TUParameters::xSetScanInfo (:75-342) arrives as per-shape tables built once by the host (capi.cu: buildDqTables).
//
// The file is plain C++ without CUDA syntax outside VVB_HD so that tests can compile the very same code for the CPU and run it against the reference
// (tests/hostbuild/).  No scaling lists, no transform skip (the reference routes TS blocks to QuantRDOQ2), luma.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define VVB_HD __host__ __device__ __forceinline__
#else
#define VVB_HD inline
#endif

namespace vvbdq {

enum { SCAN_ISCSBB = 0, SCAN_SOCSBB = 1, SCAN_EOCSBB = 2 };
enum { RICEMAX = 32, SCALE_BITS = 15, MAX_SIG_CTX = 12, MAX_GTX_CTX = 21 };

struct DqRates                       // RateEstimator (DepQuant.h:139-175) for one (TU shape, CABAC state)
{
  int32_t lastBitsX[32], lastBitsY[32];
  int32_t sigSbbBits[2][2];          // m_sigSbbFracBits[ctx].intBits[0 / 1]
  int32_t sigBits[3][MAX_SIG_CTX][2];// m_sigFracBits[set][ctx].intBits
  int32_t gtxBits[MAX_GTX_CTX][6];   // m_gtxFracBits[ctx].bits
};

struct DqQuant                       // Quantizer (DepQuant.h:190-215) after initQuantBlock
{
  int32_t qShift, maxQIdx, thresLast, distShift;
  int64_t qAdd, qScale, distAdd, distStepAdd, distOrgFact;
};

struct DqScanInfo                    // ScanInfo (DepQuant.h:82-100) without the fields that are constant per shape
{
  int16_t rasterPos, sbbPos, nextSbbRight, nextSbbBelow;
  int8_t  insidePos, nextInsidePos, spt, posX, posY, sigCtxOffsetNext, gtxCtxOffsetNext;
  uint8_t numInv, invInPos[5];
  uint8_t pad[3];
};
struct DqNbOut { uint16_t maxDist, num, outPos[5], pad; };    // NbInfoOut (DepQuant.h:69-74)

struct DqShape                       // TUParameters (DepQuant.h:103-131)
{
  int32_t width, height, numCoeff, numSbb;          // numCoeff / numSbb of the non-zero-out region (min(32, .))
  const DqScanInfo* scanInfo;                       // [numCoeff]
  const DqNbOut*    nbOut;                          // [numCoeff]
};

struct DqDec { int64_t rdCost[4]; int16_t absLevel[4]; int8_t prevId[4]; };      // Decisions (DepQuant.h:268-273)
struct DqTrellis { int16_t absLevel[4]; int8_t prevId[4]; };                      // what the backward pass needs of a Decisions record (12 bytes)

struct DqState                       // StateMem (DepQuant.h:275-307)
{
  int64_t rdCost[4];
  int16_t remRegBins[4];
  int32_t sbbBits0[4], sbbBits1[4];
  uint8_t tplAcc[16][4], sum1st[16][4], absVal[16][4];
  uint8_t ctxSig[4], ctxCff[4];
  uint8_t numSig[4];
  int8_t  refSbbCtxId[4];
  int8_t  goRicePar[4], goRiceZero[4];
  int32_t cffBitsCtxOffset;
  int32_t anyRemRegBinsLt4;
  int32_t initRemRegBins;
};

#define VVB_DQ_RDCOST_INIT ( INT64_MAX >> 1 )

#ifdef __CUDACC__
#define VVB_DQ_TAB __device__ __constant__
#else
#define VVB_DQ_TAB static const
#endif
// g_goRiceBits (DepQuant.cpp:674-680): bits of the Golomb-Rice / exp-Golomb remainder, scaled by 2^15
VVB_DQ_TAB int32_t c_goRiceBits[4][RICEMAX] = {
  {  32768,  65536,  98304, 131072, 163840, 196608, 262144, 262144, 327680, 327680, 327680, 327680, 393216, 393216, 393216, 393216, 393216, 393216, 393216, 393216, 458752, 458752, 458752, 458752, 458752, 458752, 458752, 458752, 458752, 458752, 458752, 458752 },
  {  65536,  65536,  98304,  98304, 131072, 131072, 163840, 163840, 196608, 196608, 229376, 229376, 294912, 294912, 294912, 294912, 360448, 360448, 360448, 360448, 360448, 360448, 360448, 360448, 425984, 425984, 425984, 425984, 425984, 425984, 425984, 425984 },
  {  98304,  98304,  98304,  98304, 131072, 131072, 131072, 131072, 163840, 163840, 163840, 163840, 196608, 196608, 196608, 196608, 229376, 229376, 229376, 229376, 262144, 262144, 262144, 262144, 327680, 327680, 327680, 327680, 327680, 327680, 327680, 327680 },
  { 131072, 131072, 131072, 131072, 131072, 131072, 131072, 131072, 163840, 163840, 163840, 163840, 163840, 163840, 163840, 163840, 196608, 196608, 196608, 196608, 196608, 196608, 196608, 196608, 229376, 229376, 229376, 229376, 229376, 229376, 229376, 229376 } };
// g_auiGoRiceParsCoeff (Rom.cpp:1464-1467)
VVB_DQ_TAB uint8_t c_goRicePars[32] = { 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3 };

VVB_HD int dq_min( int a, int b ) { return a < b ? a : b; }
VVB_HD int dq_max( int a, int b ) { return a > b ? a : b; }

// per-TU working set.  levels: CommonCtx memory, 8 chunks of (numSbb + numCoeff) bytes (DepQuant.h:232-245); trellis: [numCoeff][2] records
struct DqWork
{
  DqState   curr, skip, prev;
  uint8_t*  ctxMem;          // 8 * (numSbb + numCoeff) bytes
  DqTrellis* trellis;        // numCoeff * 2
  int64_t   rdCost0[4];      // m_trellis[0][0].rdCost (the only costs the path search reads)
  int       currSet;         // 0: m_currSbbCtx = m_allSbbCtx, 4: the other half (CommonCtx::swap)
};

VVB_HD uint8_t* dq_sbb_flags( const DqWork& wk, const DqShape& sh, int set, int k ) { return wk.ctxMem + ( set + k ) * ( sh.numSbb + sh.numCoeff ); }
VVB_HD uint8_t* dq_levels( const DqWork& wk, const DqShape& sh, int set, int k )    { return dq_sbb_flags( wk, sh, set, k ) + sh.numSbb; }

VVB_HD void dq_init_state( DqState& st, int k )                                     // initStates, :682-695
{
  st.rdCost[k] = VVB_DQ_RDCOST_INIT; st.ctxCff[k] = 0; st.ctxSig[k] = 0; st.numSig[k] = 0; st.refSbbCtxId[k] = -1; st.remRegBins[k] = 4;
  st.cffBitsCtxOffset = 0; st.goRicePar[k] = 0; st.goRiceZero[k] = 0; st.sbbBits0[k] = 0; st.sbbBits1[k] = 0;
}

VVB_HD const int32_t* dq_sig_bits( const DqRates& r, int stateId, int ctx ) { return r.sigBits[dq_max( stateId - 1, 0 )][ctx]; }    // sigFlagBits(stateId), DepQuant.h:149-152

// checkRdCosts, :697-775 (rrgEnsured = false form; the <true> form of checkAllRdCosts is the same arithmetic when remRegBins >= 4 holds)
VVB_HD void dq_check_rd_costs( bool rrg, int stateId, int spt, int levA, int64_t distA, int levB, int64_t distB, DqDec& dec, int idxAZ, int idxB, const DqState& st, const DqRates& r )
{
  const int32_t* goRiceTab = c_goRiceBits[st.goRicePar[stateId]];
  int64_t rdCostA = st.rdCost[stateId] + distA, rdCostB = st.rdCost[stateId] + distB, rdCostZ = st.rdCost[stateId];
  if( rrg || st.remRegBins[stateId] >= 4 )
  {
    const int32_t* cffBits = r.gtxBits[st.ctxCff[stateId]];
    const int32_t* sigBits = dq_sig_bits( r, stateId, st.ctxSig[stateId] );
    if( levA < 4 ) rdCostA += cffBits[levA];
    else { const unsigned v = (unsigned)( levA - 4 ) >> 1; rdCostA += cffBits[levA - ( v << 1 )] + goRiceTab[v < RICEMAX - 1 ? v : RICEMAX - 1]; }
    if( levB < 4 ) rdCostB += cffBits[levB];
    else { const unsigned v = (unsigned)( levB - 4 ) >> 1; rdCostB += cffBits[levB - ( v << 1 )] + goRiceTab[v < RICEMAX - 1 ? v : RICEMAX - 1]; }
    if( spt == SCAN_ISCSBB )      { rdCostA += sigBits[1]; rdCostB += sigBits[1]; rdCostZ += sigBits[0]; }
    else if( spt == SCAN_SOCSBB ) { rdCostA += st.sbbBits1[stateId] + sigBits[1]; rdCostB += st.sbbBits1[stateId] + sigBits[1]; rdCostZ += st.sbbBits1[stateId] + sigBits[0]; }
    else if( st.numSig[stateId] ) { rdCostA += sigBits[1]; rdCostB += sigBits[1]; rdCostZ += sigBits[0]; }
    else rdCostZ = VVB_DQ_RDCOST_INIT;
  }
  else
  {
    rdCostA += ( 1 << SCALE_BITS ) + goRiceTab[levA <= st.goRiceZero[stateId] ? levA - 1 : dq_min( levA, RICEMAX - 1 )];
    rdCostB += ( 1 << SCALE_BITS ) + goRiceTab[levB <= st.goRiceZero[stateId] ? levB - 1 : dq_min( levB, RICEMAX - 1 )];
    rdCostZ += goRiceTab[st.goRiceZero[stateId]];
  }
  if( rdCostA < rdCostZ && rdCostA < dec.rdCost[idxAZ] ) { dec.rdCost[idxAZ] = rdCostA; dec.absLevel[idxAZ] = (int16_t) levA; dec.prevId[idxAZ] = (int8_t) stateId; }
  else if( rdCostZ < dec.rdCost[idxAZ] )                 { dec.rdCost[idxAZ] = rdCostZ; dec.absLevel[idxAZ] = 0; dec.prevId[idxAZ] = (int8_t) stateId; }
  if( rdCostB < dec.rdCost[idxB] ) { dec.rdCost[idxB] = rdCostB; dec.absLevel[idxB] = (int16_t) levB; dec.prevId[idxB] = (int8_t) stateId; }
}

// checkRdCostsOdd1, :785-838.  cffBits1[ctx] of the reference is gtxFracBits[ctx].bits[1] (:1203-1207).
VVB_HD void dq_check_rd_costs_odd1( bool rrg, int stateId, int spt, int64_t deltaDist, DqDec& dec, int idxA, int idxZ, const DqState& st, const DqRates& r )
{
  int64_t rdCostA = st.rdCost[stateId] + deltaDist, rdCostZ = st.rdCost[stateId];
  if( rrg || st.remRegBins[stateId] >= 4 )
  {
    const int32_t* sigBits = dq_sig_bits( r, stateId, st.ctxSig[stateId] );
    rdCostA += r.gtxBits[st.ctxCff[stateId]][1];
    if( spt == SCAN_ISCSBB )      { rdCostA += sigBits[1]; rdCostZ += sigBits[0]; }
    else if( spt == SCAN_SOCSBB ) { rdCostA += st.sbbBits1[stateId] + sigBits[1]; rdCostZ += st.sbbBits1[stateId] + sigBits[0]; }
    else if( st.numSig[stateId] ) { rdCostA += sigBits[1]; rdCostZ += sigBits[0]; }
    else rdCostZ = VVB_DQ_RDCOST_INIT;
  }
  else
  {
    const int32_t* goRiceTab = c_goRiceBits[st.goRicePar[stateId]];
    rdCostA += ( 1 << SCALE_BITS ) + goRiceTab[0];
    rdCostZ += goRiceTab[st.goRiceZero[stateId]];
  }
  if( rdCostA < dec.rdCost[idxA] ) { dec.rdCost[idxA] = rdCostA; dec.absLevel[idxA] = 1; dec.prevId[idxA] = (int8_t) stateId; }
  if( rdCostZ < dec.rdCost[idxZ] ) { dec.rdCost[idxZ] = rdCostZ; dec.absLevel[idxZ] = 0; dec.prevId[idxZ] = (int8_t) stateId; }
}

// checkRdCostStart, :848-869
VVB_HD void dq_check_rd_cost_start( int32_t lastOffset, int lev, int64_t dist, DqDec& dec, int idx, const DqRates& r )
{
  const int32_t* cffBits = r.gtxBits[0];
  int64_t rdCost = dist + lastOffset;
  if( lev < 4 ) rdCost += cffBits[lev];
  else { const unsigned v = (unsigned)( lev - 4 ) >> 1; rdCost += cffBits[lev - ( v << 1 )] + c_goRiceBits[0][v < RICEMAX ? v : RICEMAX - 1]; }
  if( rdCost < dec.rdCost[idx] ) { dec.rdCost[idx] = rdCost; dec.absLevel[idx] = (int16_t) lev; dec.prevId[idx] = -1; }
}

// setRiceParam, :890-905
VVB_HD void dq_set_rice_param( int stateId, int insidePos, DqState& st, bool ge4 )
{
  if( st.remRegBins[stateId] < 4 || ge4 )
  {
    const int sumAbs = st.sum1st[insidePos][stateId];
    const int sumSub = st.remRegBins[stateId] < 4 ? 0 : 4 * 5;
    const int sumAll = dq_max( dq_min( 31, sumAbs - sumSub ), 0 );
    st.goRicePar[stateId] = (int8_t) c_goRicePars[sumAll];
    if( st.remRegBins[stateId] < 4 ) st.goRiceZero[stateId] = (int8_t)( ( stateId < 2 ? 1 : 2 ) << st.goRicePar[stateId] );       // g_auiGoRicePosCoeff0, Rom.h:137-140
  }
}

// the context part both update functions end with (:987-999, :1071-1083)
VVB_HD void dq_next_ctx( int stateId, const DqScanInfo& si, DqState& curr )
{
  if( curr.remRegBins[stateId] >= 4 )
  {
    const int sumAbs1 = curr.tplAcc[si.nextInsidePos][stateId] & 31, sumNum = curr.tplAcc[si.nextInsidePos][stateId] >> 5;
    const int sumGt1 = sumAbs1 - sumNum;
    curr.ctxSig[stateId] = (uint8_t)( si.sigCtxOffsetNext + dq_min( ( sumAbs1 + 1 ) >> 1, 3 ) );
    curr.ctxCff[stateId] = (uint8_t)( si.gtxCtxOffsetNext + dq_min( sumGt1, 4 ) );
  }
  else curr.anyRemRegBinsLt4 = 1;
}

// update1State, :907-1000
VVB_HD void dq_update1_state( int stateId, const DqScanInfo& si, const DqDec& dec, DqState& curr, const DqState& prev, bool capSum )
{
  curr.rdCost[stateId] = dec.rdCost[stateId];
  if( dec.prevId[stateId] > -2 )
  {
    const int lev = dec.absLevel[stateId];
    if( dec.prevId[stateId] >= 0 )
    {
      const int prevId = dec.prevId[stateId];
      curr.numSig[stateId] = (uint8_t)( prev.numSig[prevId] + ( lev ? 1 : 0 ) );
      curr.refSbbCtxId[stateId] = prev.refSbbCtxId[prevId];
      curr.sbbBits0[stateId] = prev.sbbBits0[prevId]; curr.sbbBits1[stateId] = prev.sbbBits1[prevId];
      curr.remRegBins[stateId] = (int16_t)( prev.remRegBins[prevId] - 1 );
      if( curr.remRegBins[stateId] >= 4 ) curr.remRegBins[stateId] = (int16_t)( curr.remRegBins[stateId] - ( lev < 2 ? lev : 3 ) );
      for( int i = 0; i < 16; i++ ) { curr.tplAcc[i][stateId] = prev.tplAcc[i][prevId]; curr.sum1st[i][stateId] = prev.sum1st[i][prevId]; curr.absVal[i][stateId] = prev.absVal[i][prevId]; }
    }
    else
    {
      curr.numSig[stateId] = 1; curr.refSbbCtxId[stateId] = -1;
      curr.remRegBins[stateId] = (int16_t)( prev.initRemRegBins - ( lev < 2 ? lev : 3 ) );
      for( int i = 0; i < 16; i++ ) { curr.tplAcc[i][stateId] = 0; curr.sum1st[i][stateId] = 0; curr.absVal[i][stateId] = 0; }
    }
    if( lev )
    {
      curr.absVal[si.insidePos][stateId] = (uint8_t) dq_min( 126 + ( lev & 1 ), lev );
      const int min4or5 = dq_min( 4 + ( lev & 1 ), lev );
      for( int k = 0; k < si.numInv && k < 5; k++ )
      {
        const int p = si.invInPos[k];
        curr.tplAcc[p][stateId] = (uint8_t)( curr.tplAcc[p][stateId] + 32 + min4or5 );
        // saturating byte add (:956-966).  What is added differs between the reference's two member sets for levels above 127: the scalar update1State adds
        // uint8_t( absLevel ) (the level modulo 256), the x86 updateStates the level capped to 126 / 127 (DepQuantX86.h:86-93, 163-166: mlvl).  The encoder
        // runs the x86 members unless started with --SIMD=SCALAR; `capSum` selects them.
        const unsigned add = capSum ? (unsigned) dq_min( 126 + ( lev & 1 ), lev ) : (unsigned)(uint8_t) lev;
        const unsigned s = (unsigned) curr.sum1st[p][stateId] + add;
        curr.sum1st[p][stateId] = (uint8_t)( s > 255u ? 255u : s );
      }
    }
    dq_next_ctx( stateId, si, curr );
  }
}

// CommonCtx::update, :473-531
VVB_HD void dq_common_ctx_update( const DqShape& sh, const DqRates& r, DqWork& wk, const DqScanInfo& si, int scanIdx, int prevId, int stateId, DqState& curr )
{
  const int prevSet = wk.currSet ^ 4;
  uint8_t* sbbFlags = dq_sbb_flags( wk, sh, wk.currSet, stateId );
  uint8_t* levels   = dq_levels( wk, sh, wk.currSet, stateId );
  const int maxDist = sh.nbOut[scanIdx - 1].maxDist, sbbSize = 16;
  const int setCp = maxDist > sbbSize ? maxDist - sbbSize : 0;
  if( prevId >= 0 )
  {
    const uint8_t* pf = dq_sbb_flags( wk, sh, prevSet, prevId ); const uint8_t* pl = dq_levels( wk, sh, prevSet, prevId );
    for( int i = 0; i < sh.numSbb; i++ ) sbbFlags[i] = pf[i];
    for( int i = 0; i < setCp; i++ ) levels[scanIdx + sbbSize + i] = pl[scanIdx + sbbSize + i];
  }
  else
  {
    for( int i = 0; i < sh.numSbb; i++ ) sbbFlags[i] = 0;
    for( int i = 0; i < setCp; i++ ) levels[scanIdx + sbbSize + i] = 0;
  }
  sbbFlags[si.sbbPos] = curr.numSig[stateId] ? 1 : 0;
  const int sigNSbb = ( ( si.nextSbbRight ? sbbFlags[si.nextSbbRight] : 0 ) || ( si.nextSbbBelow ? sbbFlags[si.nextSbbBelow] : 0 ) ) ? 1 : 0;
  curr.refSbbCtxId[stateId] = (int8_t) stateId;
  curr.sbbBits0[stateId] = r.sigSbbBits[sigNSbb][0]; curr.sbbBits1[stateId] = r.sigSbbBits[sigNSbb][1];
  if( sigNSbb || ( ( si.nextSbbRight && si.nextSbbBelow ) ? sbbFlags[si.nextSbbBelow + 1] : 0 ) )
  {
    const int scanBeg = scanIdx - sbbSize;
    const DqNbOut* nbOut = sh.nbOut + scanBeg;
    const uint8_t* absLevels = levels + scanBeg;
    for( int id = 0; id < sbbSize; id++, nbOut++ )
    {
      if( nbOut->num )
      {
        int sumAbs = 0, sumAbs1 = 0, sumNum = 0;
        for( int k = 0; k < nbOut->num && k < 5; k++ ) { const int t = absLevels[nbOut->outPos[k]]; sumAbs += t; sumAbs1 += dq_min( 4 + ( t & 1 ), t ); sumNum += t ? 1 : 0; }
        curr.tplAcc[id][stateId] = (uint8_t)( ( sumNum << 5 ) | sumAbs1 );
        curr.sum1st[id][stateId] = (uint8_t) dq_min( 255, sumAbs );
      }
    }
  }
}

// update1StateEOS, :1002-1084
VVB_HD void dq_update1_state_eos( int stateId, const DqShape& sh, const DqRates& r, DqWork& wk, const DqScanInfo& si, int scanIdx, const DqDec& dec, DqState& curr, const DqState& prev )
{
  const DqState& skip = wk.skip;
  curr.rdCost[stateId] = dec.rdCost[stateId];
  if( dec.prevId[stateId] > -2 )
  {
    const int lev = dec.absLevel[stateId];
    if( dec.prevId[stateId] >= 4 )
    {
      const int prevId = dec.prevId[stateId] - 4;
      curr.numSig[stateId] = 0; curr.remRegBins[stateId] = skip.remRegBins[prevId]; curr.refSbbCtxId[stateId] = (int8_t) prevId;
      for( int i = 0; i < 16; i++ ) curr.absVal[i][stateId] = 0;
    }
    else if( dec.prevId[stateId] >= 0 )
    {
      const int prevId = dec.prevId[stateId];
      curr.numSig[stateId] = (uint8_t)( prev.numSig[prevId] + ( lev ? 1 : 0 ) );
      curr.refSbbCtxId[stateId] = prev.refSbbCtxId[prevId];
      curr.remRegBins[stateId] = (int16_t)( prev.remRegBins[prevId] - 1 );
      if( curr.remRegBins[stateId] >= 4 ) curr.remRegBins[stateId] = (int16_t)( curr.remRegBins[stateId] - ( lev < 2 ? lev : 3 ) );
      for( int i = 0; i < 16; i++ ) curr.absVal[i][stateId] = prev.absVal[i][prevId];
    }
    else
    {
      curr.numSig[stateId] = 1; curr.refSbbCtxId[stateId] = -1;
      curr.remRegBins[stateId] = (int16_t)( prev.initRemRegBins - ( lev < 2 ? lev : 3 ) );
      for( int i = 0; i < 16; i++ ) curr.absVal[i][stateId] = 0;
    }
    curr.absVal[si.insidePos][stateId] = (uint8_t) dq_min( 126 + ( lev & 1 ), lev );
    uint8_t* lv = dq_levels( wk, sh, wk.currSet, stateId ) + scanIdx;             // getLevelPtrs, DepQuant.h:251-257
    for( int i = 0; i < 16; i++ ) { lv[i] = curr.absVal[i][stateId]; curr.tplAcc[i][stateId] = 0; curr.sum1st[i][stateId] = 0; curr.absVal[i][stateId] = 0; }
    dq_common_ctx_update( sh, r, wk, si, scanIdx, curr.refSbbCtxId[stateId], stateId, curr );
    curr.numSig[stateId] = 0;
    dq_next_ctx( stateId, si, curr );
  }
}

// xDecide, :1266-1386
VVB_HD void dq_decide( const DqQuant& q, const DqRates& r, DqWork& wk, const DqScanInfo& si, int absCoeff, int32_t lastOffset, DqDec& dec, bool zeroOut )
{
  for( int k = 0; k < 4; k++ ) { dec.rdCost[k] = VVB_DQ_RDCOST_INIT >> 2; dec.absLevel[k] = -1; dec.prevId[k] = -2; }     // startDec[0], :1113-1127
  const DqState& skip = wk.skip;
  if( zeroOut )
  {
    if( si.spt == SCAN_EOCSBB )
      for( int k = 0; k < 4; k++ ) { dec.rdCost[k] = skip.rdCost[k] + skip.sbbBits0[k]; dec.absLevel[k] = 0; dec.prevId[k] = (int8_t)( 4 | k ); }      // checkRdCostSkipSbbZeroOut
    return;
  }
  DqState& prev = wk.curr;
  const int64_t scaledOrg = (int64_t) absCoeff * q.qScale;
  int qIdx = (int)( ( scaledOrg + q.qAdd ) >> q.qShift );
  if( qIdx < 0 )
  {
    const int64_t scaledAdd = q.distStepAdd - scaledOrg * q.distOrgFact;
    const int64_t distA = ( ( scaledAdd + 0 * q.distStepAdd ) * 1 + q.distAdd ) >> q.distShift;
    const int64_t distB = ( ( scaledAdd + 1 * q.distStepAdd ) * 2 + q.distAdd ) >> q.distShift;
    const bool rrg = !prev.anyRemRegBinsLt4;       // the reference then calls m_checkAllRdCostsOdd1 = the same four calls with rrgEnsured (:840-846, 1313-1317)
    if( !rrg )
      for( int k = 0; k < 4; k++ ) dq_set_rice_param( k, si.insidePos, prev, false );
    dq_check_rd_costs_odd1( rrg, 0, si.spt, distB, dec, 2, 0, prev, r );
    dq_check_rd_costs_odd1( rrg, 1, si.spt, distB, dec, 0, 2, prev, r );
    dq_check_rd_costs_odd1( rrg, 2, si.spt, distA, dec, 3, 1, prev, r );
    dq_check_rd_costs_odd1( rrg, 3, si.spt, distA, dec, 1, 3, prev, r );
    dq_check_rd_cost_start( lastOffset, 1, distB, dec, 2, r );
  }
  else
  {
    qIdx = dq_max( 1, dq_min( q.maxQIdx, qIdx ) );
    const int64_t scaledAdd = qIdx * q.distStepAdd - scaledOrg * q.distOrgFact;
    int lev[4]; int64_t dist[4];
    for( int j = 0; j < 4; j++ )
    {
      const int slot = ( qIdx + j ) & 3;
      dist[slot] = ( ( scaledAdd + j * q.distStepAdd ) * ( qIdx + j ) + q.distAdd ) >> q.distShift;
      lev[slot]  = ( qIdx + j + 1 ) >> 1;
    }
    const bool cff02ge4 = lev[0] >= 4, cff13ge4 = lev[3] >= 4;
    if( prev.anyRemRegBinsLt4 || cff02ge4 ) { dq_set_rice_param( 0, si.insidePos, prev, cff02ge4 ); dq_set_rice_param( 1, si.insidePos, prev, cff02ge4 ); }
    if( prev.anyRemRegBinsLt4 || cff13ge4 ) { dq_set_rice_param( 2, si.insidePos, prev, cff13ge4 ); dq_set_rice_param( 3, si.insidePos, prev, cff13ge4 ); }
    const bool rrg = !( cff02ge4 || cff13ge4 || prev.anyRemRegBinsLt4 );      // m_checkAllRdCosts = checkRdCosts<true> (:777-783, 1369-1373)
    dq_check_rd_costs( rrg, 0, si.spt, lev[0], dist[0], lev[2], dist[2], dec, 0, 2, prev, r );
    dq_check_rd_costs( rrg, 1, si.spt, lev[0], dist[0], lev[2], dist[2], dec, 2, 0, prev, r );
    dq_check_rd_costs( rrg, 2, si.spt, lev[3], dist[3], lev[1], dist[1], dec, 1, 3, prev, r );
    dq_check_rd_costs( rrg, 3, si.spt, lev[3], dist[3], lev[1], dist[1], dec, 3, 1, prev, r );
    dq_check_rd_cost_start( lastOffset, lev[0], dist[0], dec, 0, r );
    dq_check_rd_cost_start( lastOffset, lev[2], dist[2], dec, 2, r );
  }
  if( si.spt == SCAN_EOCSBB )
    for( int k = 0; k < 4; k++ )                                                                                             // checkRdCostSkipSbb, :871-880
    {
      const int64_t rdCost = skip.rdCost[k] + skip.sbbBits0[k];
      if( rdCost < dec.rdCost[k] ) { dec.rdCost[k] = rdCost; dec.absLevel[k] = 0; dec.prevId[k] = (int8_t)( 4 | k ); }
    }
}

// DepQuant::xQuantDQ, :1129-1264.  coef: raster [height][width] transform coefficients; q: raster levels (written completely).
// capSum: see dq_update1_state.  zeroOutMts: the TU uses MTS / SBT transforms (effective width / height 16 for a dimension of 32, :1153-1158); lfnst: cu.lfnstIdx != 0 (:1162-1165).
VVB_HD void dq_quant_tu( const DqShape& sh, const DqQuant& q, const DqRates& r, bool zeroOutMts, bool lfnst, bool capSum, const int32_t* coef, int16_t* qOut, DqWork& wk, int32_t* absSumOut, int32_t* lastPosOut )
{
  const int W = sh.width, H = sh.height;
  for( int i = 0; i < W * H; i++ ) qOut[i] = 0;
  *absSumOut = 0;
  bool zeroOut = false;
  int effW = W, effH = H;
  if( zeroOutMts ) { effH = H == 32 ? 16 : H; effW = W == 32 ? 16 : W; zeroOut = effH < H || effW < W; }
  const bool zeroOutForThres = zeroOut || 32 < H || 32 < W;
  int firstTestPos = dq_min( W, 32 ) * dq_min( H, 32 ) - 1;
  if( lfnst ) firstTestPos = ( ( W == 4 && H == 4 ) || ( W == 8 && H == 8 ) ) ? 7 : 15;
  const int zeroOutW = ( W == 32 && zeroOut ) ? 16 : 32, zeroOutH = ( H == 32 && zeroOut ) ? 16 : 32;
  const int defaultTh = q.thresLast / (int)( q.qScale << 2 );
  for( ; firstTestPos >= 0; firstTestPos-- )                                                                                 // findFirstPos, :58-73
  {
    const DqScanInfo& si = sh.scanInfo[firstTestPos];
    if( zeroOutForThres && ( si.posX >= zeroOutW || si.posY >= zeroOutH ) ) continue;
    const int c = coef[si.rasterPos];
    if( ( c < 0 ? -c : c ) > defaultTh ) break;
  }
  if( firstTestPos < 0 ) { *lastPosOut = -1; return; }

  wk.currSet = 0;                                                                                                            // CommonCtx::reset
  for( int k = 0; k < 4; k++ ) { dq_init_state( wk.curr, k ); dq_init_state( wk.skip, k ); }
  for( int i = 0; i < 16; i++ ) for( int k = 0; k < 4; k++ ) { wk.curr.sum1st[i][k] = 0; wk.curr.tplAcc[i][k] = 0; wk.curr.absVal[i][k] = 0; }    // the reference clears sum1st only; the
                                                                                                   // other two are overwritten before a live state reads them (:1209-1211)
  const int effectW = dq_min( 32, effW ), effectH = dq_min( 32, effH );
  wk.curr.initRemRegBins = ( effectW * effectH * 28 ) / 16;                                                                  // MAX_TU_LEVEL_CTX_CODED_BIN_CONSTRAINT = 28
  wk.curr.anyRemRegBinsLt4 = 1;

  for( int scanIdx = firstTestPos; scanIdx >= 0; scanIdx-- )                                                                 // xDecideAndUpdate, :1388-1414
  {
    const DqScanInfo& si = sh.scanInfo[scanIdx];
    const int c = coef[si.rasterPos];
    const bool zo = zeroOut && ( si.posX >= effW || si.posY >= effH );
    DqDec dec;
    dq_decide( q, r, wk, si, c < 0 ? -c : c, r.lastBitsX[si.posX] + r.lastBitsY[si.posY], dec, zo );
    DqTrellis& t0 = wk.trellis[2 * scanIdx];
    for( int k = 0; k < 4; k++ ) { t0.absLevel[k] = dec.absLevel[k]; t0.prevId[k] = dec.prevId[k]; }
    if( scanIdx == 0 ) { for( int k = 0; k < 4; k++ ) wk.rdCost0[k] = dec.rdCost[k]; }
    if( scanIdx )
    {
      if( si.spt == SCAN_SOCSBB )                                                                                            // memcpy( skip, curr, StateMemSkipCpySize ): rdCost, remRegBins, sbbBits0
        for( int k = 0; k < 4; k++ ) { wk.skip.rdCost[k] = wk.curr.rdCost[k]; wk.skip.remRegBins[k] = wk.curr.remRegBins[k]; wk.skip.sbbBits0[k] = wk.curr.sbbBits0[k]; }
      if( si.insidePos == 0 )
      {
        wk.currSet ^= 4;                                                                                                     // m_commonCtx.swap()
        wk.prev = wk.curr;                                                                                                   // updateStatesEOS, :1099-1110
        wk.curr.anyRemRegBinsLt4 = 0;
        for( int k = 0; k < 4; k++ ) dq_update1_state_eos( k, sh, r, wk, si, scanIdx, dec, wk.curr, wk.prev );
        wk.curr.cffBitsCtxOffset = si.gtxCtxOffsetNext;
        wk.trellis[2 * scanIdx + 1] = t0;                                                                                    // memcpy( decisions + 1, decisions )
      }
      else if( !zo )
      {
        wk.prev = wk.curr;                                                                                                   // updateStates, :1086-1097
        wk.curr.anyRemRegBinsLt4 = 0;
        for( int k = 0; k < 4; k++ ) dq_update1_state( k, si, dec, wk.curr, wk.prev, capSum );
        wk.curr.cffBitsCtxOffset = si.gtxCtxOffsetNext;
      }
    }
  }
  // best path (:1238-1249) and backward scan (:1251-1262)
  int prevId = -1; int64_t minPathCost = 0;
  for( int k = 0; k < 4; k++ ) if( wk.rdCost0[k] < minPathCost ) { prevId = k; minPathCost = wk.rdCost0[k]; }
  int scanIdx = 0, absSum = 0;
  for( ; prevId >= 0; scanIdx++ )
  {
    // m_trellis[.][1] is written at the first position of a coefficient group only; everywhere else it holds startDec[1] = { level 0, prevId 4 | k }
    // (:1121-1126, 1436-1439): a skipped group is walked with zero levels up to the group's first position
    if( prevId >= 4 && ( scanIdx & 15 ) ) continue;
    const DqTrellis& t = wk.trellis[2 * scanIdx + ( prevId >> 2 )];
    const int absLevel = t.absLevel[prevId & 3];
    const int blkpos = sh.scanInfo[scanIdx].rasterPos;
    qOut[blkpos] = (int16_t)( coef[blkpos] < 0 ? -absLevel : absLevel );
    absSum += absLevel;
    prevId = t.prevId[prevId & 3];
  }
  *absSumOut = absSum; *lastPosOut = scanIdx - 1;
}

} // namespace vvbdq
