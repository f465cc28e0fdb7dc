// rdoq_core.h -- the fast rate-distortion optimised quantiser of one TU: QuantRDOQ2::xRateDistOptQuantFast<bSBH, false> (CommonLib/QuantRDOQ2.cpp:475-1281) with
// its helpers xiGetICRateCost (:320-401), xiGetCostLast (:445-461), _dist (:468-473) and the CoeffCodingContext members it drives (ContextModelling.h:158-269,
// ContextModelling.cpp:113-133), restated for one thread per TU.  This is what Quant::m_RDOQ == 2 selects (the presets faster and fast, vvencCfg.cpp:2675, 2737)
// for every TU that is not transform skipped, and what DepQuant::quant falls back to in slices without dependent quantisation (DepQuant.cpp:1486-1489).
//
// What stays on the host: everything that depends on the encoder's entropy-coding state.  The fractional bits of the contexts the routine reads (significance
// set 0, parity, greater-1, greater-2, significant-group, the last-position tables of xInitLastPosBitsTab :408-434, the coded-block-flag context of :1185-1226)
// arrive as `RqRates`; the per-call constants (quantiser scale and shift, the error scale of xSetErrScaleCoeffNoScalingList :203-219, thresholds) as `RqPar`
// (rdoq_host.h).  No scaling lists, no transform skip (rateDistOptQuantTS), sides 4..64 (coefficient groups are always 4x4 there, g_log2SbbSize).
//
// The template bookkeeping of the reference (m_tplBuf: per position the sum of min( 4 + ( l & 1 ), l ) and the count of the five already-coded neighbours,
// kept in step with the level buffer by absVal1stPass / remAbsVal1stPass at every change) is a pure function of the level buffer, so it is read from the
// levels directly (the form sigCtxIdAbs, ContextModelling.h:115-156, uses).
//
// Plain C++ without CUDA syntax outside VVB_HD: the test suite compiles the very same text for the CPU (g++), where it is pinned against the reference's member.
#pragma once
#include <stdint.h>

#ifndef VVB_HD
#ifdef __CUDACC__
#define VVB_HD __host__ __device__ __forceinline__
#else
#define VVB_HD inline
#endif
#endif

namespace vvbrq {

enum { RQ_SCALE_BITS = 15, RQ_ERR_SCALE_SHIFT = 20 /* COEFF_ERR_SCALE_PRECISION_BITS, QuantRDOQ2.cpp:84 */, RQ_SBH_THRESHOLD = 4, RQ_REMAIN_BIN_REDUCTION = 5 };

typedef int64_t cost_t;              // QuantRDOQ2.h:59

struct RqRates                       // BinFracBits::intBits of the contexts, as FracBitsAccess::getFracBitsArray returns them at the point of the call
{
  int32_t sigBits[12][2];            // Ctx::SigFlag[chType]( ctxOfs ): set 0 (state 0), ctxOfs 0..11 (luma) / 0..7 (chroma)
  int32_t parBits[21][2];            // Ctx::ParFlag[chType]( ctxOffsetAbs )
  int32_t gt1Bits[21][2];            // Ctx::GtxFlag[chType + 2]( ctxOffsetAbs )  (greater1CtxIdAbs, ContextModelling.h:239)
  int32_t gt2Bits[21][2];            // Ctx::GtxFlag[chType]( ctxOffsetAbs )      (greater2CtxIdAbs, :240)
  int32_t sigGroupBits[2][2];        // Ctx::SigCoeffGroup[chType]( sigRight | sigLower )
  int32_t lastBitsX[16], lastBitsY[16];   // m_lastBitsX / m_lastBitsY[chType][ctxId] after xInitLastPosBitsTab
  int32_t cbfBits[2];                // the coded-block-flag context of :1185-1226 (QtRootCbf for inter luma, QtCbf otherwise); zeros when the flag is inferred
  int32_t pad[2];
};

struct RqPar
{
  int32_t width, height, log2W;      // TU size
  int32_t regionW;                   // min( 32, width ): row pitch of the scan table entries
  int32_t numCG;                     // iCGNum, :553
  int32_t firstScanPos;              // the position the search for the first non-zero coefficient starts from, :554-559
  int32_t quantScale;                // g_quantScales[needsSqrt2][rem], :518
  int32_t errScale;                  // xGetErrScaleCoeffNoScalingList, :519
  int32_t qBits;                     // iQBits, :522
  int32_t useThres;                  // thres / ( quantScale << 2 ), :573-583
  int32_t remRegBins;                // ( tbAreaAfterCoefZeroOut * 28 ) >> 4, :539
  int32_t signHiding;                // bSBH
  int32_t isChroma;                  // channel type of the component (context offsets)
  int32_t pad;
  double  lambda;                    // Quant::m_dLambda
};

#ifdef __CUDACC__
#define VVB_RQ_TAB __device__ __constant__
#else
#define VVB_RQ_TAB static const
#endif
VVB_RQ_TAB uint8_t c_rqGoRicePars[32] = { 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3 };      // g_auiGoRiceParsCoeff, Rom.cpp:1464-1467
VVB_RQ_TAB uint8_t c_rqGroupIdx[32]   = { 0, 1, 2, 3, 4, 4, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7, 8, 8, 8, 8, 8, 8, 8, 8, 9, 9, 9, 9, 9, 9, 9, 9 };      // g_uiGroupIdx, Rom.cpp:1463 (positions < 32)

VVB_HD int rq_min( int a, int b ) { return a < b ? a : b; }
VVB_HD int rq_max( int a, int b ) { return a > b ? a : b; }
VVB_HD int rq_abs( int a ) { return a < 0 ? -a : a; }

VVB_HD cost_t rq_icost( const RqPar& P, int rate ) { return (cost_t)( P.lambda * rate ); }                   // xiGetICost, :303-306

// xiGetICRateCost, :320-401
VVB_HD cost_t rq_level_rate_cost( const RqPar& P, uint32_t absLevel, const int32_t* par, const int32_t* gt1, const int32_t* gt2, int remRegBins, uint32_t goRiceZero, uint32_t goRice )
{
  cost_t rate = (cost_t) 1 << RQ_SCALE_BITS;                    // xGetIEPRate: the sign bin
  if( remRegBins < 4 )
  {
    uint32_t symbol = ( absLevel == 0 ? goRiceZero : absLevel <= goRiceZero ? absLevel - 1 : absLevel );
    uint32_t length;
    const int threshold = RQ_REMAIN_BIN_REDUCTION;
    if( symbol < ( (uint32_t) threshold << goRice ) )
    {
      length = symbol >> goRice;
      rate += (cost_t)( length + 1 + goRice ) << RQ_SCALE_BITS;
    }
    else
    {
      length = goRice;
      symbol = symbol - ( (uint32_t) threshold << goRice );
      while( symbol >= ( 1u << length ) ) symbol -= ( 1u << ( length++ ) );
      rate += (cost_t)( threshold + length + 1 - goRice + length ) << RQ_SCALE_BITS;
    }
  }
  else
  {
    const uint32_t cthres = 4;
    if( absLevel >= cthres )
    {
      uint32_t symbol = ( absLevel - cthres ) >> 1;
      uint32_t length;
      const int threshold = RQ_REMAIN_BIN_REDUCTION;
      if( symbol < ( (uint32_t) threshold << goRice ) )
      {
        length = symbol >> goRice;
        rate += (cost_t)( length + 1 + goRice ) << RQ_SCALE_BITS;
      }
      else
      {
        length = goRice;
        symbol = symbol - ( (uint32_t) threshold << goRice );
        while( symbol >= ( 1u << length ) ) symbol -= ( 1u << ( length++ ) );
        rate += (cost_t)( threshold + length + 1 - goRice + length ) << RQ_SCALE_BITS;
      }
      rate += gt1[1];
      rate += par[( absLevel - 2 ) & 1];
      rate += gt2[1];
    }
    else if( absLevel == 1 ) { rate += gt1[0]; }
    else if( absLevel == 2 ) { rate += gt1[1]; rate += par[0]; rate += gt2[0]; }
    else if( absLevel == 3 ) { rate += gt1[1]; rate += par[1]; rate += gt2[0]; }
    else rate = 0;
  }
  return rq_icost( P, (int) rate );
}

VVB_HD cost_t rq_dist( cost_t err, cost_t errScale )              // _dist, :468-473
{
  const int64_t s = ( err * errScale ) >> RQ_ERR_SCALE_SHIFT;
  return s * s;
}

// the five already-coded neighbours of (x, y) in the level buffer (levels are kept as magnitudes until the sign pass at the end)
#define VVB_RQ_TEMPLATE( q, W, H, x, y, OP ) { const int16_t* pd_ = ( q ) + ( y ) * ( W ) + ( x ); \
  if( ( x ) < ( W ) - 1 ) { OP( pd_[1] ); if( ( x ) < ( W ) - 2 ) OP( pd_[2] ); if( ( y ) < ( H ) - 1 ) OP( pd_[( W ) + 1] ); } \
  if( ( y ) < ( H ) - 1 ) { OP( pd_[( W )] ); if( ( y ) < ( H ) - 2 ) OP( pd_[2 * ( W )] ); } }

// one TU.  scan: scan position -> raster index inside the scanned region (row pitch P.regionW), grouped 4x4 diagonal scan; coef [h][w] TCoeff; q [h][w] levels (written);
// absSum / lastPos as the reference leaves uiAbsSum / tu.lastPos (lastPos -1 where the reference does not write it: nothing coded)
VVB_HD void rq_quant_tu( const RqPar& P, const RqRates& R, const int32_t* scan, const int32_t* coef, int16_t* q, int32_t* absSumOut, int32_t* lastPosOut )
{
  const int W = P.width, H = P.height, lw = P.log2W;
  const int lrw = ( P.regionW == 32 ? 5 : P.regionW == 16 ? 4 : P.regionW == 8 ? 3 : 2 );
  const bool bSBH = P.signHiding != 0, luma = P.isChroma == 0;
  const int iQBits = P.qBits, quantScale = P.quantScale;
  const int iQOffset = 1 << ( iQBits - 1 );
  const cost_t iErrScale = P.errScale;
  const int widthInGroups = rq_min( 32, W ) >> 2, heightInGroups = rq_min( 32, H ) >> 2;
#define RQ_BLKPOS( sp ) ( ( ( scan[sp] >> lrw ) << lw ) + ( scan[sp] & ( P.regionW - 1 ) ) )

  for( int i = 0; i < W * H; i++ ) q[i] = 0;                      // :513

  cost_t piCostCoeff[16], piCostSig[16], piCostCoeff0[16], piCostDeltaSBH[16];
  int    piAddSBH[16];
  for( int i = 0; i < 16; i++ ) { piCostCoeff[i] = 0; piCostSig[i] = 0; piCostCoeff0[i] = 0; piCostDeltaSBH[i] = 0; piAddSBH[i] = 0; }

  cost_t iCodedCostBlock = 0, iUncodedCostBlock = 0;
  int    iLastScanPos = -1, lastSubSetId = -1;
  bool   lastOptFinished = false;
  cost_t bestTotalCost = INT64_MAX / 2;
  int    remRegBins = P.remRegBins;
  uint32_t goRiceParam = 0;
  int    uiAbsSum = 0;
  const int iCGSize = 16, iCGSizeM1 = 15, log2CGSize = 4;
  uint64_t sigGroupFlags = 0;                                     // m_sigCoeffGroupFlag, indexed by the raster position of the group
  int    tmplCpDiag = -1, tmplCpSum1 = -1;                        // CoeffCodingContext::m_tmplCpDiag / m_tmplCpSum1 (persist from position to position)

  int iScanPos = P.firstScanPos;
  for( ; iScanPos > 0; iScanPos-- ) if( coef[RQ_BLKPOS( iScanPos )] ) break;        // :561-567

  int subSetId = iScanPos >> log2CGSize;
  for( ; subSetId >= 0; subSetId-- )
  {
    int    iNZbeforePos0 = 0, uiAbsSumCG = 0;
    cost_t iCodedCostCG = 0, iUncodedCostCG = 0;
    int    iScanPosinCG = iScanPos & ( iCGSize - 1 );

    if( iLastScanPos < 0 && iScanPos >= 16 )                      // :599-656 (the SIMD and the scalar form test the same positions: everything above iScanPos is zero)
    {
      bool allSmaller = true;
      for( int xp = iScanPosinCG, xs = iScanPos; allSmaller && xp >= 0; xp--, xs-- ) allSmaller &= rq_abs( coef[RQ_BLKPOS( xs )] ) <= P.useThres;
      if( allSmaller ) { iScanPos -= iScanPosinCG + 1; continue; }
    }

    // group position and the context of its significant-group flag (initSubblock, ContextModelling.cpp:113-133)
    const int cgRaster = scan[subSetId << 4], cgX = ( cgRaster & ( P.regionW - 1 ) ) >> 2, cgY = ( cgRaster >> lrw ) >> 2;
    const int subSetPos = cgY * widthInGroups + cgX;
    const uint64_t cgBit = (uint64_t) 1 << subSetPos;
    int remRegBinsStartCG = remRegBins;
    int sigGroupCtx = 0;

    bool findLast = iLastScanPos < 0;
    for( ;; )
    {
      if( findLast )                                              // findlast2, :658-686
      {
        for( ; iScanPosinCG >= 0; iScanPosinCG--, iScanPos-- )
        {
          const uint32_t maxAbsLevel = (uint32_t)( ( rq_abs( coef[RQ_BLKPOS( iScanPos )] ) * quantScale + iQOffset ) >> iQBits );
          if( maxAbsLevel ) { iLastScanPos = iScanPos; lastSubSetId = subSetId; break; }
        }
        findLast = false;
      }
      {
        const unsigned sigRight = ( cgX + 1 ) < widthInGroups  ? (unsigned)( ( sigGroupFlags >> ( subSetPos + 1 ) ) & 1 ) : 0u;
        const unsigned sigLower = ( cgY + 1 ) < heightInGroups ? (unsigned)( ( sigGroupFlags >> ( subSetPos + widthInGroups ) ) & 1 ) : 0u;
        sigGroupCtx = (int)( sigRight | sigLower );
      }
      remRegBinsStartCG = remRegBins;

      bool again = false;
      for( ; iScanPosinCG >= 0; iScanPosinCG--, iScanPos-- )      // :697-969
      {
        const int raster = scan[iScanPos], posX = raster & ( P.regionW - 1 ), posY = raster >> lrw;
        const int uiBlkPos = ( posY << lw ) + posX;
        const int iScaledLevel = rq_abs( coef[uiBlkPos] ) * quantScale;
        const int iAbsLevel = ( iScaledLevel + iQOffset ) >> iQBits;

        int ctxIdSig = 0;
        if( iScanPos != iLastScanPos )                            // sigCtxIdAbsWithAcc( iScanPos, 0 ), ContextModelling.h:158-178
        {
          int numPos = 0, sumAbs = 0;
#define RQ_UPD( v ) { const int a_ = ( v ); sumAbs += rq_min( 4 + ( a_ & 1 ), a_ ); numPos += a_ != 0; }
          VVB_RQ_TEMPLATE( q, W, H, posX, posY, RQ_UPD )
#undef RQ_UPD
          const int diag = posX + posY;
          ctxIdSig = rq_min( ( sumAbs + 1 ) >> 1, 3 ) + ( diag < 2 ? 4 : 0 );
          if( luma ) ctxIdSig += diag < 5 ? 4 : 0;
          tmplCpDiag = diag; tmplCpSum1 = sumAbs - numPos;
        }
        int ctxOffset = 0;                                        // ctxOffsetAbs, ContextModelling.h:227-236
        if( tmplCpDiag != -1 )
        {
          ctxOffset  = rq_min( tmplCpSum1, 4 ) + 1;
          ctxOffset += ( !tmplCpDiag ? ( luma ? 15 : 5 ) : luma ? ( tmplCpDiag < 3 ? 10 : ( tmplCpDiag < 10 ? 5 : 0 ) ) : 0 );
        }
        const int32_t* fbPar = R.parBits[ctxOffset];
        const int32_t* fbGt1 = R.gt1Bits[ctxOffset];
        const int32_t* fbGt2 = R.gt2Bits[ctxOffset];
        const int32_t* fbSig = R.sigBits[ctxIdSig];
        uint32_t goRiceZero = 0;

        if( remRegBins < 4 )                                      // :731-736
        {
          int sum = 0;
#define RQ_SUM( v ) { sum += ( v ); }
          VVB_RQ_TEMPLATE( q, W, H, posX, posY, RQ_SUM )
#undef RQ_SUM
          const int sumAbs = rq_max( rq_min( sum, 31 ), 0 );      // templateAbsSum( ., ., 0 )
          goRiceParam = c_rqGoRicePars[sumAbs];
          goRiceZero  = 1u << goRiceParam;                        // g_auiGoRicePosCoeff0( 0, . ), Rom.h:137-140
        }

        piCostCoeff0[iScanPosinCG] = rq_dist( iScaledLevel, iErrScale );

        uint32_t uiLevel = 0;
        if( iAbsLevel == 0 )                                      // :748-770
        {
          piCostSig  [iScanPosinCG] = rq_icost( P, fbSig[0] );
          piCostCoeff[iScanPosinCG] = piCostCoeff0[iScanPosinCG] + piCostSig[iScanPosinCG];
          if( bSBH )
          {
            const cost_t iErr1  = iScaledLevel - ( (int64_t) 1 << iQBits );
            const cost_t iDist1 = rq_dist( iErr1, iErrScale );
            const cost_t iRate1 = remRegBins < 4 ? rq_level_rate_cost( P, 1, fbPar, fbGt1, fbGt2, remRegBins, goRiceZero, goRiceParam ) -
                                                   rq_level_rate_cost( P, 0, fbPar, fbGt1, fbGt2, remRegBins, goRiceZero, goRiceParam )
                                                 : (cost_t) fbGt1[0];
            const cost_t iCost1 = iDist1 + iRate1 + rq_icost( P, fbSig[1] );
            piCostDeltaSBH[iScanPosinCG] = iCost1 - piCostCoeff[iScanPosinCG];
            piAddSBH      [iScanPosinCG] = 1;
          }
        }
        else
        {
          const int iFloor = (int)( iScaledLevel >> iQBits );
          const int iCeil  = iFloor + 1;

          if( remRegBins >= 4 && iScanPos != iLastScanPos && iCeil >= 4 )     // :777-781
          {
            int sum = 0;
#define RQ_SUM( v ) { sum += ( v ); }
            VVB_RQ_TEMPLATE( q, W, H, posX, posY, RQ_SUM )
#undef RQ_SUM
            goRiceParam = c_rqGoRicePars[rq_max( rq_min( sum - 5 * 4, 31 ), 0 )];
          }

          if( iScanPos == iLastScanPos )                          // last level, :783-835
          {
            piCostSig[iScanPosinCG] = 0;
            cost_t iCurrCostF = piCostCoeff0[iScanPosinCG];
            if( iFloor )
            {
              const cost_t iErrF = iScaledLevel - ( iFloor << iQBits );
              iCurrCostF = rq_dist( iErrF, iErrScale ) + rq_level_rate_cost( P, iFloor, fbPar, fbGt1, fbGt2, remRegBins, goRiceZero, goRiceParam );
            }
            const cost_t iErrC = iScaledLevel - ( iCeil << iQBits );
            const cost_t iCurrCostC = rq_dist( iErrC, iErrScale ) + rq_level_rate_cost( P, iCeil, fbPar, fbGt1, fbGt2, remRegBins, goRiceZero, goRiceParam );

            if( iCurrCostC < iCurrCostF )
            {
              uiLevel = iCeil;
              piCostCoeff[iScanPosinCG] = iCurrCostC;
              if( bSBH ) { piCostDeltaSBH[iScanPosinCG] = iCurrCostF - iCurrCostC; piAddSBH[iScanPosinCG] = -1; }
            }
            else
            {
              if( iFloor == 0 )                                   // the candidate last position quantises to zero: look for the next one (goto findlast2, :816-827)
              {
                iLastScanPos = -1; lastSubSetId = -1;
                iScanPos--; iScanPosinCG--;
                again = true;
                break;
              }
              uiLevel = iFloor;
              piCostCoeff[iScanPosinCG] = iCurrCostF;
              if( bSBH ) { piCostDeltaSBH[iScanPosinCG] = iCurrCostC - iCurrCostF; piAddSBH[iScanPosinCG] = 1; }
            }
          }
          else
          {
            const cost_t iCostSig1 = rq_icost( P, fbSig[1] );
            if( iCeil < 3 )                                       // levels 0, 1, 2, :840-907
            {
              const cost_t iCostSig0 = rq_icost( P, fbSig[0] );
              cost_t iBestCost = piCostCoeff0[iScanPosinCG] + iCostSig0;
              cost_t iBestCostSig = iCostSig0;
              cost_t iCostF = iBestCost;
              uiLevel = 0;
              if( iFloor == 1 )
              {
                const cost_t iErrF = iScaledLevel - ( iFloor << iQBits );
                iCostF = rq_dist( iErrF, iErrScale ) + iCostSig1 + rq_level_rate_cost( P, iFloor, fbPar, fbGt1, fbGt2, remRegBins, goRiceZero, goRiceParam );
                if( iCostF < iBestCost )
                {
                  uiLevel = iFloor; iBestCost = iCostF; iBestCostSig = iCostSig1;
                  if( bSBH ) { piCostDeltaSBH[iScanPosinCG] = iBestCost - iCostF; piAddSBH[iScanPosinCG] = -1; }
                }
                else
                {
                  if( bSBH ) { piCostDeltaSBH[iScanPosinCG] = iCostF - iBestCost; piAddSBH[iScanPosinCG] = 1; }
                }
              }
              const cost_t iErrC = iScaledLevel - ( iCeil << iQBits );
              const cost_t iCostC = rq_dist( iErrC, iErrScale ) + iCostSig1 + rq_level_rate_cost( P, iCeil, fbPar, fbGt1, fbGt2, remRegBins, goRiceZero, goRiceParam );
              if( iCostC < iBestCost )
              {
                uiLevel = iCeil;
                piCostCoeff[iScanPosinCG] = iCostC;
                piCostSig[iScanPosinCG]   = iCostSig1;
                if( bSBH ) { piCostDeltaSBH[iScanPosinCG] = iCostF - iCostC; piAddSBH[iScanPosinCG] = -1; }
              }
              else
              {
                piCostCoeff[iScanPosinCG] = iBestCost;
                piCostSig[iScanPosinCG]   = iBestCostSig;
                if( bSBH ) { piCostDeltaSBH[iScanPosinCG] = iCostC - iCostF; piAddSBH[iScanPosinCG] = 1; }
              }
            }
            else                                                  // levels x, x + 1, :908-940
            {
              const cost_t iErrF = iScaledLevel - ( iFloor << iQBits );
              const cost_t iCostF = rq_dist( iErrF, iErrScale ) + iCostSig1 + rq_level_rate_cost( P, iFloor, fbPar, fbGt1, fbGt2, remRegBins, goRiceZero, goRiceParam );
              const cost_t iErrC = iScaledLevel - ( iCeil << iQBits );
              const cost_t iCostC = rq_dist( iErrC, iErrScale ) + iCostSig1 + rq_level_rate_cost( P, iCeil, fbPar, fbGt1, fbGt2, remRegBins, goRiceZero, goRiceParam );
              piCostSig[iScanPosinCG] = iCostSig1;
              if( iCostC < iCostF )
              {
                uiLevel = iCeil;
                piCostCoeff[iScanPosinCG] = iCostC;
                if( bSBH ) { piCostDeltaSBH[iScanPosinCG] = iCostF - iCostC; piAddSBH[iScanPosinCG] = -1; }
              }
              else
              {
                uiLevel = iFloor;
                piCostCoeff[iScanPosinCG] = iCostF;
                if( bSBH ) { piCostDeltaSBH[iScanPosinCG] = iCostC - iCostF; piAddSBH[iScanPosinCG] = 1; }
              }
            }
          }
          q[uiBlkPos] = (int16_t) uiLevel;                        // :942
          if( uiLevel )
          {
            uiAbsSumCG    += uiLevel;
            iNZbeforePos0 += iScanPosinCG;
            sigGroupFlags |= cgBit;                               // setSigGroup
          }
        }

        if( ( ( iScanPos & iCGSizeM1 ) == 0 ) && ( iScanPos > 0 ) ) goRiceParam = 0;                      // :956-963
        else if( remRegBins >= 4 ) remRegBins -= ( uiLevel < 2 ? (int) uiLevel : 3 ) + ( iScanPos != iLastScanPos );

        iUncodedCostCG += piCostCoeff0[iScanPosinCG];
        iCodedCostCG   += piCostCoeff[iScanPosinCG];
      }
      if( !again ) break;
      findLast = true;
    }

    //================== group significance flag, :971-1036 ===================
    cost_t iCostCoeffGroupSig = 0;
    if( lastSubSetId >= 0 )
    {
      if( subSetId )
      {
        const cost_t iCostCoeffGroupSig0 = rq_icost( P, R.sigGroupBits[sigGroupCtx][0] );
        if( !( sigGroupFlags & cgBit ) )
        {
          iCodedCostCG = iUncodedCostCG + iCostCoeffGroupSig0;
          iCostCoeffGroupSig = iCostCoeffGroupSig0;
        }
        else
        {
          if( subSetId < lastSubSetId )
          {
            const cost_t iCostCoeffGroupSig1 = rq_icost( P, R.sigGroupBits[sigGroupCtx][1] );
            iCostCoeffGroupSig = iCostCoeffGroupSig1;
            if( !iNZbeforePos0 ) iCodedCostCG -= piCostSig[0];
            const cost_t iUncodedCostCGTmp = iUncodedCostCG + iCostCoeffGroupSig0;
            iCodedCostCG += iCostCoeffGroupSig1;
            if( iUncodedCostCGTmp < iCodedCostCG )                // cheaper as an all-zero group
            {
              sigGroupFlags &= ~cgBit;                            // resetSigGroup
              iCodedCostCG = iUncodedCostCGTmp;
              iCostCoeffGroupSig = iCostCoeffGroupSig0;
              remRegBins = remRegBinsStartCG;
              for( int p = iCGSize - 1; p >= 0; p-- ) q[RQ_BLKPOS( subSetId * iCGSize + p )] = 0;
              uiAbsSumCG = 0;
              if( lastSubSetId == subSetId ) { iCodedCostCG = 0; iUncodedCostCG = 0; iLastScanPos = -1; lastSubSetId = -1; }
            }
          }
          else sigGroupFlags |= cgBit;
        }
      }
    }

    //===== last position cost, :1038-1095 =====
    bestTotalCost += iCodedCostCG;
    if( !lastOptFinished )
    {
      if( sigGroupFlags & cgBit )
      {
        cost_t codedCostBlockTmp = iUncodedCostBlock + iCodedCostCG - iCostCoeffGroupSig;
        const int startPosInCG = subSetId == lastSubSetId ? iLastScanPos % iCGSize : iCGSizeM1;
        int newAbsSumCG = uiAbsSumCG;
        int bestLastIdxP1 = iLastScanPos + 1;
        for( int pc = startPosInCG; pc >= 0; pc-- )
        {
          const int sp = ( subSetId << log2CGSize ) + pc;
          const int raster = scan[sp], px = raster & ( P.regionW - 1 ), py = raster >> lrw;
          const int bp = ( py << lw ) + px;
          if( q[bp] )
          {
            // xiGetCostLast, :445-461
            const uint32_t ctxX = c_rqGroupIdx[px], ctxY = c_rqGroupIdx[py];
            uint32_t uiCost = (uint32_t) R.lastBitsX[ctxX] + (uint32_t) R.lastBitsY[ctxY];
            if( ctxX > 3 ) uiCost += ( 1u << RQ_SCALE_BITS ) * ( ( ctxX - 2 ) >> 1 );
            if( ctxY > 3 ) uiCost += ( 1u << RQ_SCALE_BITS ) * ( ( ctxY - 2 ) >> 1 );
            const cost_t iCostLast = rq_icost( P, (int) uiCost );
            const cost_t totalCost = codedCostBlockTmp + iCostLast - piCostSig[pc];
            if( totalCost < bestTotalCost )
            {
              bestLastIdxP1 = sp + 1; bestTotalCost = totalCost; lastSubSetId = subSetId; uiAbsSumCG = newAbsSumCG; uiAbsSum = 0;
            }
            if( q[bp] > 1 ) { lastOptFinished = true; break; }
            newAbsSumCG -= 1;
            codedCostBlockTmp -= piCostCoeff[pc];
            codedCostBlockTmp += piCostCoeff0[pc];
          }
          else codedCostBlockTmp -= piCostSig[pc];
        }
        for( int sp = bestLastIdxP1; sp <= iLastScanPos; sp++ ) q[RQ_BLKPOS( sp )] = 0;
        iLastScanPos = bestLastIdxP1 - 1;
      }
    }

    //=============== sign bit hiding, :1097-1167 ================
    if( bSBH )
    {
      if( uiAbsSumCG >= 2 )
      {
        const int iSubPos = subSetId * iCGSize;
        int iLastNZPosInCG = -1, iFirstNZPosInCG = iCGSize;
        for( int n = 0; n < iCGSize; n++ ) if( q[RQ_BLKPOS( n + iSubPos )] ) { iFirstNZPosInCG = n; break; }
        if( lastSubSetId == subSetId )
        {
          iLastNZPosInCG = iLastScanPos % iCGSize;
          if( q[RQ_BLKPOS( iLastScanPos )] == 1 && piAddSBH[iLastNZPosInCG] == -1 ) piCostDeltaSBH[iLastNZPosInCG] -= ( 4 << RQ_SCALE_BITS );
        }
        else
        {
          for( int n = iCGSize - 1; n >= 0; n-- ) if( q[RQ_BLKPOS( n + iSubPos )] ) { iLastNZPosInCG = n; break; }
        }
        if( iLastNZPosInCG - iFirstNZPosInCG >= RQ_SBH_THRESHOLD )
        {
          iCodedCostCG -= rq_icost( P, 1 << RQ_SCALE_BITS );
          const bool bSign = coef[RQ_BLKPOS( iSubPos + iFirstNZPosInCG )] < 0;
          if( (int) bSign != ( uiAbsSumCG & 0x1 ) )
          {
            const int iLastPosInCG = ( lastSubSetId == subSetId ) ? iLastNZPosInCG : iCGSize - 1;
            int64_t iMinCostDelta = INT64_MAX;
            int iMinCostPos = -1;
            if( q[RQ_BLKPOS( iFirstNZPosInCG + iSubPos )] > 1 ) { iMinCostDelta = piCostDeltaSBH[iFirstNZPosInCG]; iMinCostPos = iFirstNZPosInCG; }
            for( int n = 0; n < iFirstNZPosInCG; n++ )
              if( ( coef[RQ_BLKPOS( iSubPos + n )] < 0 ) == bSign )
                if( piCostDeltaSBH[n] < iMinCostDelta ) { iMinCostDelta = piCostDeltaSBH[n]; iMinCostPos = n; }
            for( int n = iFirstNZPosInCG + 1; n <= iLastPosInCG; n++ )
              if( piCostDeltaSBH[n] < iMinCostDelta ) { iMinCostDelta = piCostDeltaSBH[n]; iMinCostPos = n; }
            const int bp = RQ_BLKPOS( iMinCostPos + iSubPos );
            q[bp] = (int16_t)( q[bp] + piAddSBH[iMinCostPos] );
            uiAbsSumCG   += piAddSBH[iMinCostPos];
            iCodedCostCG += iMinCostDelta;
          }
        }
      }
    }

    iCodedCostBlock   += iCodedCostCG;
    iUncodedCostBlock += iUncodedCostCG;
    uiAbsSum += uiAbsSumCG;
  }

  iCodedCostBlock = bestTotalCost;                                // :1177

  if( iLastScanPos < 0 ) { *absSumOut = uiAbsSum; *lastPosOut = -1; return; }         // :1179-1183 (uiAbsSum is 0 there)

  iUncodedCostBlock += rq_icost( P, R.cbfBits[0] );               // :1185-1226 (the caller resolved which context applies; zeros when the flag is inferred)
  iCodedCostBlock   += rq_icost( P, R.cbfBits[1] );

  if( iUncodedCostBlock <= iCodedCostBlock )                      // :1228-1233
  {
    for( int i = 0; i < W * H; i++ ) q[i] = 0;
    *absSumOut = 0; *lastPosOut = -1;
    return;
  }
  if( bSBH && q[RQ_BLKPOS( iLastScanPos )] == 0 )                 // :1237-1249
  {
    int sp = iLastScanPos - 1;
    for( ; sp >= 0; sp-- ) if( q[RQ_BLKPOS( sp )] ) break;
    iLastScanPos = sp;
  }
  for( int sp = 0; sp <= iLastScanPos; sp++ )                     // signs, :1251-1257
  {
    const int bp = RQ_BLKPOS( sp );
    const int level = q[bp];
    const int iSign = coef[bp] >> 31;
    q[bp] = (int16_t)( ( iSign ^ level ) - iSign );
  }
  *absSumOut = uiAbsSum; *lastPosOut = iLastScanPos;
#undef RQ_BLKPOS
}

// Second engine of the same routine (vvb_set_rdoq_engine 2): identical decisions, fewer instructions per coefficient.
//  * The template of a position is not gathered from its five neighbours when the position is visited (a quarter of the executed instructions of the first engine,
//    profiles/r02zz_src_rdoq_kernel_32x32.txt): it is ACCUMULATED, as the reference does in m_tplBuf (absVal1stPass / remAbsVal1stPass, ContextModelling.h:180-225), in the
//    level slot of the position itself -- a position that has not been visited yet holds no level, so its slot carries ( count << 5 | sum ) of the decided neighbours; when a
//    level is set, changed (sign-bit hiding) or cleared (group zero-out, last-position optimisation), the five positions to the left / above that are still unvisited are
//    updated.  A position is unvisited iff its coefficient group comes earlier in the scan than the group being worked on (cgIdx: group raster position -> group scan index).
//  * lambda * bits of the significance flags and of the levels 1..3 with context-coded bins come from tables computed once per call (RqCost, same double product).
struct RqCost
{
  int64_t sig[12][2];                // xiGetICost( sigBits[ctx][bin] )
  int64_t lvl[21][3];                // xiGetICRateCost( 1 / 2 / 3, ... ) with remRegBins >= 4 for greater-1 / parity / greater-2 context offset ctx
};
#define VVB_RQ_ENC( L ) ( ( L ) ? 32 + rq_min( 4 + ( ( L ) & 1 ), ( L ) ) : 0 )
// add `delta` to the accumulators of the unvisited dependents of (x, y): every dependent when ALL is set (the position itself is being visited: everything to its left /
// above is still ahead), else only those in groups that come earlier in the scan than group `curCG`
#define VVB_RQ_DEPS( q, W, x, y, delta, ALL, cgIdx, wInGroups, curCG ) { int16_t* pq_ = ( q ) + ( y ) * ( W ) + ( x ); \
  if( ( y ) > 1 && ( ( ALL ) || cgIdx[( ( ( y ) - 2 ) >> 2 ) * ( wInGroups ) + ( ( x ) >> 2 )] < ( curCG ) ) ) pq_[-2 * ( W )] = (int16_t)( pq_[-2 * ( W )] + ( delta ) ); \
  if( ( y ) > 0 && ( x ) > 0 && ( ( ALL ) || cgIdx[( ( ( y ) - 1 ) >> 2 ) * ( wInGroups ) + ( ( ( x ) - 1 ) >> 2 )] < ( curCG ) ) ) pq_[-( W ) - 1] = (int16_t)( pq_[-( W ) - 1] + ( delta ) ); \
  if( ( y ) > 0 && ( ( ALL ) || cgIdx[( ( ( y ) - 1 ) >> 2 ) * ( wInGroups ) + ( ( x ) >> 2 )] < ( curCG ) ) ) pq_[-( W )] = (int16_t)( pq_[-( W )] + ( delta ) ); \
  if( ( x ) > 1 && ( ( ALL ) || cgIdx[( ( y ) >> 2 ) * ( wInGroups ) + ( ( ( x ) - 2 ) >> 2 )] < ( curCG ) ) ) pq_[-2] = (int16_t)( pq_[-2] + ( delta ) ); \
  if( ( x ) > 0 && ( ( ALL ) || cgIdx[( ( y ) >> 2 ) * ( wInGroups ) + ( ( ( x ) - 1 ) >> 2 )] < ( curCG ) ) ) pq_[-1] = (int16_t)( pq_[-1] + ( delta ) ); }
VVB_HD void rq_quant_tu_v2( const RqPar& P, const RqRates& R, const RqCost& C, const int32_t* scan, const uint8_t* cgIdx, const int32_t* coef, int16_t* q, int32_t* absSumOut, int32_t* lastPosOut )
{
  const int W = P.width, H = P.height, lw = P.log2W;
  const int lrw = ( P.regionW == 32 ? 5 : P.regionW == 16 ? 4 : P.regionW == 8 ? 3 : 2 );
  const bool bSBH = P.signHiding != 0, luma = P.isChroma == 0;
  const int iQBits = P.qBits, quantScale = P.quantScale;
  const int iQOffset = 1 << ( iQBits - 1 );
  const cost_t iErrScale = P.errScale;
  const int widthInGroups = rq_min( 32, W ) >> 2, heightInGroups = rq_min( 32, H ) >> 2;
#define RQ_BLKPOS( sp ) ( ( ( scan[sp] >> lrw ) << lw ) + ( scan[sp] & ( P.regionW - 1 ) ) )

  for( int i = 0; i < W * H; i++ ) q[i] = 0;                      // :513

#define RQ_LVL_COST( L, par_, gt1_, gt2_, rrb_, grz_, grp_ ) ( ( ( rrb_ ) >= 4 && (uint32_t)( L ) - 1u < 3u ) ? C.lvl[ctxOffset][(uint32_t)( L ) - 1u] : rq_level_rate_cost( P, ( L ), par_, gt1_, gt2_, rrb_, grz_, grp_ ) )
  cost_t piCostCoeff[16], piCostSig[16], piCostCoeff0[16], piCostDeltaSBH[16];
  int    piAddSBH[16];
  for( int i = 0; i < 16; i++ ) { piCostCoeff[i] = 0; piCostSig[i] = 0; piCostCoeff0[i] = 0; piCostDeltaSBH[i] = 0; piAddSBH[i] = 0; }

  cost_t iCodedCostBlock = 0, iUncodedCostBlock = 0;
  int    iLastScanPos = -1, lastSubSetId = -1;
  bool   lastOptFinished = false;
  cost_t bestTotalCost = INT64_MAX / 2;
  int    remRegBins = P.remRegBins;
  uint32_t goRiceParam = 0;
  int    uiAbsSum = 0;
  const int iCGSize = 16, iCGSizeM1 = 15, log2CGSize = 4;
  uint64_t sigGroupFlags = 0;                                     // m_sigCoeffGroupFlag, indexed by the raster position of the group
  int    tmplCpDiag = -1, tmplCpSum1 = -1;                        // CoeffCodingContext::m_tmplCpDiag / m_tmplCpSum1 (persist from position to position)

  int iScanPos = P.firstScanPos;
  for( ; iScanPos > 0; iScanPos-- ) if( coef[RQ_BLKPOS( iScanPos )] ) break;        // :561-567

  int subSetId = iScanPos >> log2CGSize;
  for( ; subSetId >= 0; subSetId-- )
  {
    int    iNZbeforePos0 = 0, uiAbsSumCG = 0;
    cost_t iCodedCostCG = 0, iUncodedCostCG = 0;
    int    iScanPosinCG = iScanPos & ( iCGSize - 1 );

    if( iLastScanPos < 0 && iScanPos >= 16 )                      // :599-656 (the SIMD and the scalar form test the same positions: everything above iScanPos is zero)
    {
      bool allSmaller = true;
      for( int xp = iScanPosinCG, xs = iScanPos; allSmaller && xp >= 0; xp--, xs-- ) allSmaller &= rq_abs( coef[RQ_BLKPOS( xs )] ) <= P.useThres;
      if( allSmaller ) { iScanPos -= iScanPosinCG + 1; continue; }
    }

    // group position and the context of its significant-group flag (initSubblock, ContextModelling.cpp:113-133)
    const int cgRaster = scan[subSetId << 4], cgX = ( cgRaster & ( P.regionW - 1 ) ) >> 2, cgY = ( cgRaster >> lrw ) >> 2;
    const int subSetPos = cgY * widthInGroups + cgX;
    const uint64_t cgBit = (uint64_t) 1 << subSetPos;
    int remRegBinsStartCG = remRegBins;
    int sigGroupCtx = 0;

    bool findLast = iLastScanPos < 0;
    for( ;; )
    {
      if( findLast )                                              // findlast2, :658-686
      {
        for( ; iScanPosinCG >= 0; iScanPosinCG--, iScanPos-- )
        {
          const uint32_t maxAbsLevel = (uint32_t)( ( rq_abs( coef[RQ_BLKPOS( iScanPos )] ) * quantScale + iQOffset ) >> iQBits );
          if( maxAbsLevel ) { iLastScanPos = iScanPos; lastSubSetId = subSetId; break; }
        }
        findLast = false;
      }
      {
        const unsigned sigRight = ( cgX + 1 ) < widthInGroups  ? (unsigned)( ( sigGroupFlags >> ( subSetPos + 1 ) ) & 1 ) : 0u;
        const unsigned sigLower = ( cgY + 1 ) < heightInGroups ? (unsigned)( ( sigGroupFlags >> ( subSetPos + widthInGroups ) ) & 1 ) : 0u;
        sigGroupCtx = (int)( sigRight | sigLower );
      }
      remRegBinsStartCG = remRegBins;

      bool again = false;
      for( ; iScanPosinCG >= 0; iScanPosinCG--, iScanPos-- )      // :697-969
      {
        const int raster = scan[iScanPos], posX = raster & ( P.regionW - 1 ), posY = raster >> lrw;
        const int uiBlkPos = ( posY << lw ) + posX;
        const int iScaledLevel = rq_abs( coef[uiBlkPos] ) * quantScale;
        const int iAbsLevel = ( iScaledLevel + iQOffset ) >> iQBits;

        int ctxIdSig = 0;
        if( iScanPos != iLastScanPos )                            // sigCtxIdAbsWithAcc( iScanPos, 0 ), ContextModelling.h:158-178
        {
          const int acc = q[uiBlkPos];                                // the accumulator the decided neighbours left in this (still unvisited) slot
          const int numPos = acc >> 5, sumAbs = acc & 31;
          const int diag = posX + posY;
          ctxIdSig = rq_min( ( sumAbs + 1 ) >> 1, 3 ) + ( diag < 2 ? 4 : 0 );
          if( luma ) ctxIdSig += diag < 5 ? 4 : 0;
          tmplCpDiag = diag; tmplCpSum1 = sumAbs - numPos;
        }
        int ctxOffset = 0;                                        // ctxOffsetAbs, ContextModelling.h:227-236
        if( tmplCpDiag != -1 )
        {
          ctxOffset  = rq_min( tmplCpSum1, 4 ) + 1;
          ctxOffset += ( !tmplCpDiag ? ( luma ? 15 : 5 ) : luma ? ( tmplCpDiag < 3 ? 10 : ( tmplCpDiag < 10 ? 5 : 0 ) ) : 0 );
        }
        const int32_t* fbPar = R.parBits[ctxOffset];
        const int32_t* fbGt1 = R.gt1Bits[ctxOffset];
        const int32_t* fbGt2 = R.gt2Bits[ctxOffset];
        uint32_t goRiceZero = 0;

        if( remRegBins < 4 )                                      // :731-736
        {
          int sum = 0;
#define RQ_SUM( v ) { sum += ( v ); }
          VVB_RQ_TEMPLATE( q, W, H, posX, posY, RQ_SUM )
#undef RQ_SUM
          const int sumAbs = rq_max( rq_min( sum, 31 ), 0 );      // templateAbsSum( ., ., 0 )
          goRiceParam = c_rqGoRicePars[sumAbs];
          goRiceZero  = 1u << goRiceParam;                        // g_auiGoRicePosCoeff0( 0, . ), Rom.h:137-140
        }

        piCostCoeff0[iScanPosinCG] = rq_dist( iScaledLevel, iErrScale );

        uint32_t uiLevel = 0;
        if( iAbsLevel == 0 )                                      // :748-770
        {
          piCostSig  [iScanPosinCG] = C.sig[ctxIdSig][0];
          piCostCoeff[iScanPosinCG] = piCostCoeff0[iScanPosinCG] + piCostSig[iScanPosinCG];
          if( bSBH )
          {
            const cost_t iErr1  = iScaledLevel - ( (int64_t) 1 << iQBits );
            const cost_t iDist1 = rq_dist( iErr1, iErrScale );
            const cost_t iRate1 = remRegBins < 4 ? RQ_LVL_COST( 1, fbPar, fbGt1, fbGt2, remRegBins, goRiceZero, goRiceParam ) -
                                                   RQ_LVL_COST( 0, fbPar, fbGt1, fbGt2, remRegBins, goRiceZero, goRiceParam )
                                                 : (cost_t) fbGt1[0];
            const cost_t iCost1 = iDist1 + iRate1 + C.sig[ctxIdSig][1];
            piCostDeltaSBH[iScanPosinCG] = iCost1 - piCostCoeff[iScanPosinCG];
            piAddSBH      [iScanPosinCG] = 1;
          }
        }
        else
        {
          const int iFloor = (int)( iScaledLevel >> iQBits );
          const int iCeil  = iFloor + 1;

          if( remRegBins >= 4 && iScanPos != iLastScanPos && iCeil >= 4 )     // :777-781
          {
            int sum = 0;
#define RQ_SUM( v ) { sum += ( v ); }
            VVB_RQ_TEMPLATE( q, W, H, posX, posY, RQ_SUM )
#undef RQ_SUM
            goRiceParam = c_rqGoRicePars[rq_max( rq_min( sum - 5 * 4, 31 ), 0 )];
          }

          if( iScanPos == iLastScanPos )                          // last level, :783-835
          {
            piCostSig[iScanPosinCG] = 0;
            cost_t iCurrCostF = piCostCoeff0[iScanPosinCG];
            if( iFloor )
            {
              const cost_t iErrF = iScaledLevel - ( iFloor << iQBits );
              iCurrCostF = rq_dist( iErrF, iErrScale ) + RQ_LVL_COST( iFloor, fbPar, fbGt1, fbGt2, remRegBins, goRiceZero, goRiceParam );
            }
            const cost_t iErrC = iScaledLevel - ( iCeil << iQBits );
            const cost_t iCurrCostC = rq_dist( iErrC, iErrScale ) + RQ_LVL_COST( iCeil, fbPar, fbGt1, fbGt2, remRegBins, goRiceZero, goRiceParam );

            if( iCurrCostC < iCurrCostF )
            {
              uiLevel = iCeil;
              piCostCoeff[iScanPosinCG] = iCurrCostC;
              if( bSBH ) { piCostDeltaSBH[iScanPosinCG] = iCurrCostF - iCurrCostC; piAddSBH[iScanPosinCG] = -1; }
            }
            else
            {
              if( iFloor == 0 )                                   // the candidate last position quantises to zero: look for the next one (goto findlast2, :816-827)
              {
                iLastScanPos = -1; lastSubSetId = -1;
                iScanPos--; iScanPosinCG--;
                again = true;
                break;
              }
              uiLevel = iFloor;
              piCostCoeff[iScanPosinCG] = iCurrCostF;
              if( bSBH ) { piCostDeltaSBH[iScanPosinCG] = iCurrCostC - iCurrCostF; piAddSBH[iScanPosinCG] = 1; }
            }
          }
          else
          {
            const cost_t iCostSig1 = C.sig[ctxIdSig][1];
            if( iCeil < 3 )                                       // levels 0, 1, 2, :840-907
            {
              const cost_t iCostSig0 = C.sig[ctxIdSig][0];
              cost_t iBestCost = piCostCoeff0[iScanPosinCG] + iCostSig0;
              cost_t iBestCostSig = iCostSig0;
              cost_t iCostF = iBestCost;
              uiLevel = 0;
              if( iFloor == 1 )
              {
                const cost_t iErrF = iScaledLevel - ( iFloor << iQBits );
                iCostF = rq_dist( iErrF, iErrScale ) + iCostSig1 + RQ_LVL_COST( iFloor, fbPar, fbGt1, fbGt2, remRegBins, goRiceZero, goRiceParam );
                if( iCostF < iBestCost )
                {
                  uiLevel = iFloor; iBestCost = iCostF; iBestCostSig = iCostSig1;
                  if( bSBH ) { piCostDeltaSBH[iScanPosinCG] = iBestCost - iCostF; piAddSBH[iScanPosinCG] = -1; }
                }
                else
                {
                  if( bSBH ) { piCostDeltaSBH[iScanPosinCG] = iCostF - iBestCost; piAddSBH[iScanPosinCG] = 1; }
                }
              }
              const cost_t iErrC = iScaledLevel - ( iCeil << iQBits );
              const cost_t iCostC = rq_dist( iErrC, iErrScale ) + iCostSig1 + RQ_LVL_COST( iCeil, fbPar, fbGt1, fbGt2, remRegBins, goRiceZero, goRiceParam );
              if( iCostC < iBestCost )
              {
                uiLevel = iCeil;
                piCostCoeff[iScanPosinCG] = iCostC;
                piCostSig[iScanPosinCG]   = iCostSig1;
                if( bSBH ) { piCostDeltaSBH[iScanPosinCG] = iCostF - iCostC; piAddSBH[iScanPosinCG] = -1; }
              }
              else
              {
                piCostCoeff[iScanPosinCG] = iBestCost;
                piCostSig[iScanPosinCG]   = iBestCostSig;
                if( bSBH ) { piCostDeltaSBH[iScanPosinCG] = iCostC - iCostF; piAddSBH[iScanPosinCG] = 1; }
              }
            }
            else                                                  // levels x, x + 1, :908-940
            {
              const cost_t iErrF = iScaledLevel - ( iFloor << iQBits );
              const cost_t iCostF = rq_dist( iErrF, iErrScale ) + iCostSig1 + RQ_LVL_COST( iFloor, fbPar, fbGt1, fbGt2, remRegBins, goRiceZero, goRiceParam );
              const cost_t iErrC = iScaledLevel - ( iCeil << iQBits );
              const cost_t iCostC = rq_dist( iErrC, iErrScale ) + iCostSig1 + RQ_LVL_COST( iCeil, fbPar, fbGt1, fbGt2, remRegBins, goRiceZero, goRiceParam );
              piCostSig[iScanPosinCG] = iCostSig1;
              if( iCostC < iCostF )
              {
                uiLevel = iCeil;
                piCostCoeff[iScanPosinCG] = iCostC;
                if( bSBH ) { piCostDeltaSBH[iScanPosinCG] = iCostF - iCostC; piAddSBH[iScanPosinCG] = -1; }
              }
              else
              {
                uiLevel = iFloor;
                piCostCoeff[iScanPosinCG] = iCostF;
                if( bSBH ) { piCostDeltaSBH[iScanPosinCG] = iCostC - iCostF; piAddSBH[iScanPosinCG] = 1; }
              }
            }
          }
          if( uiLevel )
          {
            uiAbsSumCG    += uiLevel;
            iNZbeforePos0 += iScanPosinCG;
            sigGroupFlags |= cgBit;                               // setSigGroup
            const int enc_ = VVB_RQ_ENC( (int) uiLevel );         // absVal1stPass
            VVB_RQ_DEPS( q, W, posX, posY, enc_, true, cgIdx, widthInGroups, subSetId )
          }
        }
        q[uiBlkPos] = (int16_t) uiLevel;                          // :942; also takes the accumulator out of a slot that stays zero

        if( ( ( iScanPos & iCGSizeM1 ) == 0 ) && ( iScanPos > 0 ) ) goRiceParam = 0;                      // :956-963
        else if( remRegBins >= 4 ) remRegBins -= ( uiLevel < 2 ? (int) uiLevel : 3 ) + ( iScanPos != iLastScanPos );

        iUncodedCostCG += piCostCoeff0[iScanPosinCG];
        iCodedCostCG   += piCostCoeff[iScanPosinCG];
      }
      if( !again ) break;
      findLast = true;
    }

    //================== group significance flag, :971-1036 ===================
    cost_t iCostCoeffGroupSig = 0;
    if( lastSubSetId >= 0 )
    {
      if( subSetId )
      {
        const cost_t iCostCoeffGroupSig0 = rq_icost( P, R.sigGroupBits[sigGroupCtx][0] );
        if( !( sigGroupFlags & cgBit ) )
        {
          iCodedCostCG = iUncodedCostCG + iCostCoeffGroupSig0;
          iCostCoeffGroupSig = iCostCoeffGroupSig0;
        }
        else
        {
          if( subSetId < lastSubSetId )
          {
            const cost_t iCostCoeffGroupSig1 = rq_icost( P, R.sigGroupBits[sigGroupCtx][1] );
            iCostCoeffGroupSig = iCostCoeffGroupSig1;
            if( !iNZbeforePos0 ) iCodedCostCG -= piCostSig[0];
            const cost_t iUncodedCostCGTmp = iUncodedCostCG + iCostCoeffGroupSig0;
            iCodedCostCG += iCostCoeffGroupSig1;
            if( iUncodedCostCGTmp < iCodedCostCG )                // cheaper as an all-zero group
            {
              sigGroupFlags &= ~cgBit;                            // resetSigGroup
              iCodedCostCG = iUncodedCostCGTmp;
              iCostCoeffGroupSig = iCostCoeffGroupSig0;
              remRegBins = remRegBinsStartCG;
              for( int p = iCGSize - 1; p >= 0; p-- )
              {
                const int rs_ = scan[subSetId * iCGSize + p], px_ = rs_ & ( P.regionW - 1 ), py_ = rs_ >> lrw;
                const int bp_ = ( py_ << lw ) + px_;
                if( q[bp_] ) { const int enc_ = -VVB_RQ_ENC( (int) q[bp_] ); VVB_RQ_DEPS( q, W, px_, py_, enc_, false, cgIdx, widthInGroups, subSetId ) q[bp_] = 0; }      // remAbsVal1stPass
              }
              uiAbsSumCG = 0;
              if( lastSubSetId == subSetId ) { iCodedCostCG = 0; iUncodedCostCG = 0; iLastScanPos = -1; lastSubSetId = -1; }
            }
          }
          else sigGroupFlags |= cgBit;
        }
      }
    }

    //===== last position cost, :1038-1095 =====
    bestTotalCost += iCodedCostCG;
    if( !lastOptFinished )
    {
      if( sigGroupFlags & cgBit )
      {
        cost_t codedCostBlockTmp = iUncodedCostBlock + iCodedCostCG - iCostCoeffGroupSig;
        const int startPosInCG = subSetId == lastSubSetId ? iLastScanPos % iCGSize : iCGSizeM1;
        int newAbsSumCG = uiAbsSumCG;
        int bestLastIdxP1 = iLastScanPos + 1;
        for( int pc = startPosInCG; pc >= 0; pc-- )
        {
          const int sp = ( subSetId << log2CGSize ) + pc;
          const int raster = scan[sp], px = raster & ( P.regionW - 1 ), py = raster >> lrw;
          const int bp = ( py << lw ) + px;
          if( q[bp] )
          {
            // xiGetCostLast, :445-461
            const uint32_t ctxX = c_rqGroupIdx[px], ctxY = c_rqGroupIdx[py];
            uint32_t uiCost = (uint32_t) R.lastBitsX[ctxX] + (uint32_t) R.lastBitsY[ctxY];
            if( ctxX > 3 ) uiCost += ( 1u << RQ_SCALE_BITS ) * ( ( ctxX - 2 ) >> 1 );
            if( ctxY > 3 ) uiCost += ( 1u << RQ_SCALE_BITS ) * ( ( ctxY - 2 ) >> 1 );
            const cost_t iCostLast = rq_icost( P, (int) uiCost );
            const cost_t totalCost = codedCostBlockTmp + iCostLast - piCostSig[pc];
            if( totalCost < bestTotalCost )
            {
              bestLastIdxP1 = sp + 1; bestTotalCost = totalCost; lastSubSetId = subSetId; uiAbsSumCG = newAbsSumCG; uiAbsSum = 0;
            }
            if( q[bp] > 1 ) { lastOptFinished = true; break; }
            newAbsSumCG -= 1;
            codedCostBlockTmp -= piCostCoeff[pc];
            codedCostBlockTmp += piCostCoeff0[pc];
          }
          else codedCostBlockTmp -= piCostSig[pc];
        }
        for( int sp = bestLastIdxP1; sp <= iLastScanPos; sp++ )
        {
          const int rs_ = scan[sp], px_ = rs_ & ( P.regionW - 1 ), py_ = rs_ >> lrw;
          const int bp_ = ( py_ << lw ) + px_;
          if( q[bp_] ) { const int enc_ = -VVB_RQ_ENC( (int) q[bp_] ); VVB_RQ_DEPS( q, W, px_, py_, enc_, false, cgIdx, widthInGroups, subSetId ) q[bp_] = 0; }
        }
        iLastScanPos = bestLastIdxP1 - 1;
      }
    }

    //=============== sign bit hiding, :1097-1167 ================
    if( bSBH )
    {
      if( uiAbsSumCG >= 2 )
      {
        const int iSubPos = subSetId * iCGSize;
        int iLastNZPosInCG = -1, iFirstNZPosInCG = iCGSize;
        for( int n = 0; n < iCGSize; n++ ) if( q[RQ_BLKPOS( n + iSubPos )] ) { iFirstNZPosInCG = n; break; }
        if( lastSubSetId == subSetId )
        {
          iLastNZPosInCG = iLastScanPos % iCGSize;
          if( q[RQ_BLKPOS( iLastScanPos )] == 1 && piAddSBH[iLastNZPosInCG] == -1 ) piCostDeltaSBH[iLastNZPosInCG] -= ( 4 << RQ_SCALE_BITS );
        }
        else
        {
          for( int n = iCGSize - 1; n >= 0; n-- ) if( q[RQ_BLKPOS( n + iSubPos )] ) { iLastNZPosInCG = n; break; }
        }
        if( iLastNZPosInCG - iFirstNZPosInCG >= RQ_SBH_THRESHOLD )
        {
          iCodedCostCG -= rq_icost( P, 1 << RQ_SCALE_BITS );
          const bool bSign = coef[RQ_BLKPOS( iSubPos + iFirstNZPosInCG )] < 0;
          if( (int) bSign != ( uiAbsSumCG & 0x1 ) )
          {
            const int iLastPosInCG = ( lastSubSetId == subSetId ) ? iLastNZPosInCG : iCGSize - 1;
            int64_t iMinCostDelta = INT64_MAX;
            int iMinCostPos = -1;
            if( q[RQ_BLKPOS( iFirstNZPosInCG + iSubPos )] > 1 ) { iMinCostDelta = piCostDeltaSBH[iFirstNZPosInCG]; iMinCostPos = iFirstNZPosInCG; }
            for( int n = 0; n < iFirstNZPosInCG; n++ )
              if( ( coef[RQ_BLKPOS( iSubPos + n )] < 0 ) == bSign )
                if( piCostDeltaSBH[n] < iMinCostDelta ) { iMinCostDelta = piCostDeltaSBH[n]; iMinCostPos = n; }
            for( int n = iFirstNZPosInCG + 1; n <= iLastPosInCG; n++ )
              if( piCostDeltaSBH[n] < iMinCostDelta ) { iMinCostDelta = piCostDeltaSBH[n]; iMinCostPos = n; }
            const int rs_ = scan[iMinCostPos + iSubPos], px_ = rs_ & ( P.regionW - 1 ), py_ = rs_ >> lrw;
            const int bp = ( py_ << lw ) + px_;
            const int encDelta_ = VVB_RQ_ENC( (int) q[bp] + piAddSBH[iMinCostPos] ) - VVB_RQ_ENC( (int) q[bp] );
            if( encDelta_ ) VVB_RQ_DEPS( q, W, px_, py_, encDelta_, false, cgIdx, widthInGroups, subSetId )
            q[bp] = (int16_t)( q[bp] + piAddSBH[iMinCostPos] );
            uiAbsSumCG   += piAddSBH[iMinCostPos];
            iCodedCostCG += iMinCostDelta;
          }
        }
      }
    }

    iCodedCostBlock   += iCodedCostCG;
    iUncodedCostBlock += iUncodedCostCG;
    uiAbsSum += uiAbsSumCG;
  }

  iCodedCostBlock = bestTotalCost;                                // :1177

  if( iLastScanPos < 0 ) { *absSumOut = uiAbsSum; *lastPosOut = -1; return; }         // :1179-1183 (uiAbsSum is 0 there)

  iUncodedCostBlock += rq_icost( P, R.cbfBits[0] );               // :1185-1226 (the caller resolved which context applies; zeros when the flag is inferred)
  iCodedCostBlock   += rq_icost( P, R.cbfBits[1] );

  if( iUncodedCostBlock <= iCodedCostBlock )                      // :1228-1233
  {
    for( int i = 0; i < W * H; i++ ) q[i] = 0;
    *absSumOut = 0; *lastPosOut = -1;
    return;
  }
  if( bSBH && q[RQ_BLKPOS( iLastScanPos )] == 0 )                 // :1237-1249
  {
    int sp = iLastScanPos - 1;
    for( ; sp >= 0; sp-- ) if( q[RQ_BLKPOS( sp )] ) break;
    iLastScanPos = sp;
  }
  for( int sp = 0; sp <= iLastScanPos; sp++ )                     // signs, :1251-1257
  {
    const int bp = RQ_BLKPOS( sp );
    const int level = q[bp];
    const int iSign = coef[bp] >> 31;
    q[bp] = (int16_t)( ( iSign ^ level ) - iSign );
  }
  *absSumOut = uiAbsSum; *lastPosOut = iLastScanPos;
#undef RQ_BLKPOS
#undef RQ_LVL_COST
}


// ------------------------------------------------------------------------------------------------------------------------------------------------------------------
// Transform-skip residual coding: QuantRDOQ::rateDistOptQuantTS (CommonLib/QuantRDOQ.cpp:1124-1336) with xGetCodedLevelTSPred (:1578-1661), xGetICRateTS (:1663-1807) and
// the transform-skip members of CoeffCodingContext (ContextModelling.h:271-407, ContextModelling.cpp:130-132) -- what QuantRDOQ2::quant runs for a transform-skipped TU
// without BDPCM (QuantRDOQ2.cpp:275-285) when Quant::m_useRDOQTS is set.  The scan runs FORWARD (group 0 first, position 0 first), the context of a position comes from its
// left and upper neighbours, costs are doubles (distortion = err * err * errorScale, rate = lambda * bits) summed in the reference's order.  The per-position cost arrays
// of the reference (m_pdCostCoeff, m_pdCostSig, m_pdCostCoeff0, m_pdCostCoeffGroupSig) are only ever read at the position that has just been written, so scalars stand in.
struct RqTsRates                     // BinFracBits::intBits of the transform-skip context sets (Contexts.cpp:821-868)
{
  int32_t sigBits[3][2];             // Ctx::TsSigFlag( numPos ), numPos = number of non-zero left / upper neighbours
  int32_t parBits[2];                // Ctx::TsParFlag( 0 )
  int32_t gtxBits[5][2];             // Ctx::TsGtxFlag( cutoffVal >> 1 ): entries 1..4 are read
  int32_t lrg1Bits[4][2];            // Ctx::TsLrg1Flag( numPos )
  int32_t signBits[6][2];            // Ctx::TsResidualSign( signCtx )
  int32_t sigGroupBits[3][2];        // Ctx::TsSigCoeffGroup( sigLeft + sigAbove )
};                                   // 44 int32

struct RqTsPar
{
  int32_t width, height, log2W;
  int32_t quantScale;                // g_quantScales[0][ qp.rem( true ) ], :1160
  int32_t qBits;                     // QUANT_SHIFT + qp.per( true ), :1159 (no transform shift, no sqrt(2) compensation)
  int32_t maxCtxBins;                // ( w * h * 7 ) >> 2, :1183
  int32_t pad[2];
  double  errorScale;                // xGetErrScaleCoeff( false, w, h, rem, 15, bitDepth, true ), QuantRDOQ.cpp:319-329
  double  lambda;                    // Quant::m_dLambda
};

VVB_HD int rq_golomb_bits( uint32_t symbol, uint32_t ricePar )            // the Golomb-Rice / exp-Golomb length of xGetICRateTS, in whole bits
{
  uint32_t length;
  const uint32_t threshold = RQ_REMAIN_BIN_REDUCTION;
  if( symbol < ( threshold << ricePar ) ) { length = symbol >> ricePar; return (int)( length + 1 + ricePar ); }
  length = ricePar;
  symbol = symbol - ( threshold << ricePar );
  while( symbol >= ( 1u << length ) ) symbol -= ( 1u << ( length++ ) );
  return (int)( threshold + length + 1 - ricePar + length );
}

// xGetICRateTS, :1663-1807
VVB_HD int rq_ts_level_rate( const RqTsRates& R, uint32_t absLevel, int remRegBins, const int32_t* fbSign, const int32_t* fbGt1, int& numCtxBins, int sign, uint32_t ricePar )
{
  if( remRegBins < 4 )                                            // everything by-pass coded
  {
    int rate = absLevel ? ( 1 << RQ_SCALE_BITS ) : 0;
    rate += rq_golomb_bits( absLevel, ricePar ) << RQ_SCALE_BITS;
    return rate;
  }
  else if( remRegBins < 8 )                                       // first pass context coded, the rest by-pass
  {
    int rate = fbSign[sign];
    if( absLevel ) numCtxBins++;
    if( absLevel > 1 )
    {
      rate += fbGt1[1];
      rate += R.parBits[( absLevel - 2 ) & 1];
      numCtxBins += 2;
      rate += rq_golomb_bits( ( absLevel - 2 ) >> 1, ricePar ) << RQ_SCALE_BITS;
    }
    else if( absLevel == 1 ) { rate += fbGt1[0]; numCtxBins++; }
    else rate = 0;
    return rate;
  }
  int rate = fbSign[sign];
  if( absLevel ) numCtxBins++;
  if( absLevel > 1 )
  {
    rate += fbGt1[1];
    rate += R.parBits[( absLevel - 2 ) & 1];
    numCtxBins += 2;
    uint32_t cutoffVal = 2;
    for( int i = 0; i < 4; i++ )
    {
      if( absLevel >= cutoffVal )
      {
        rate += R.gtxBits[cutoffVal >> 1][absLevel >= ( cutoffVal + 2 ) ? 1 : 0];
        numCtxBins++;
      }
      cutoffVal += 2;
    }
    if( absLevel >= cutoffVal ) rate += rq_golomb_bits( ( absLevel - cutoffVal ) >> 1, ricePar ) << RQ_SCALE_BITS;
  }
  else if( absLevel == 1 ) { rate += fbGt1[0]; numCtxBins++; }
  else rate = 0;
  return rate;
}

// deriveModCoeff( right, below, absCoeff, 0 ), ContextModelling.h:363-386
VVB_HD int rq_ts_mod_coeff( int rightPixel, int belowPixel, int absCoeff )
{
  if( absCoeff == 0 ) return 0;
  const int pred1 = rq_max( rq_abs( belowPixel ), rq_abs( rightPixel ) );
  if( absCoeff == pred1 ) return 1;
  return absCoeff < pred1 ? absCoeff + 1 : absCoeff;
}

// one transform-skipped TU.  scan as for rq_quant_tu; coef [h][w]: the residual as xTransformSkip copies it; q [h][w] levels (signed, written); absSum as the reference leaves it
VVB_HD void rq_ts_quant_tu( const RqTsPar& P, const RqTsRates& R, const int32_t* scan, const int32_t* coef, int16_t* q, int32_t* absSumOut )
{
  const int W = P.width, H = P.height, lw = P.log2W;
  const int regionW = rq_min( 32, W );
  const int lrw = ( regionW == 32 ? 5 : regionW == 16 ? 4 : regionW == 8 ? 3 : 2 );
  const int qBits = P.qBits;
  const int widthInGroups = W >> 2, heightInGroups = H >> 2;
  const int sbNum = ( W * H ) >> 4;
  const uint32_t entropyCodingMaximum = ( 1u << 15 ) - 1;
  uint64_t sigGroupFlags = 0;
  bool anySigCG = false;
  int remRegBins = P.maxCtxBins;
  int absSum = 0;

  for( int i = 0; i < W * H; i++ ) q[i] = 0;                      // the caller's level buffer starts cleared (TrQuant::transformNxN works on a cleared TU, and neighbours ahead in the scan read as zero)

  for( int sbId = 0; sbId < sbNum; sbId++ )
  {
    // initSubblock: group position, the context of its significant-group flag from the left and upper groups (ContextModelling.cpp:113-133)
    const int cgRaster = scan[sbId << 4], cgX = ( cgRaster & ( regionW - 1 ) ) >> 2, cgY = ( cgRaster >> lrw ) >> 2;
    const int subSetPos = cgY * widthInGroups + cgX;
    const uint64_t cgBit = (uint64_t) 1 << subSetPos;
    const int sigLeft  = cgX > 0 ? (int)( ( sigGroupFlags >> ( subSetPos - 1 ) ) & 1 ) : 0;
    const int sigAbove = cgY > 0 ? (int)( ( sigGroupFlags >> ( subSetPos - widthInGroups ) ) & 1 ) : 0;
    const int32_t* fbSigGroup = R.sigGroupBits[sigLeft + sigAbove];
    (void) heightInGroups;

    int noCoeffCoded = 0;
    double baseCost = 0.0;
    double d64CodedLevelandDist = 0.0, d64UncodedDist = 0.0, d64SigCost = 0.0;      // coeffGroupRDStats
    int iNumSbbCtxBins = 0;

    for( int scanPosInSB = 0; scanPosInSB <= 15; scanPosInSB++ )
    {
      const int scanPos = ( sbId << 4 ) + scanPosInSB;
      const int raster = scan[scanPos], posX = raster & ( regionW - 1 ), posY = raster >> lrw;
      const int blkPos = ( posY << lw ) + posX;

      const int64_t tmpLevel = (int64_t) rq_abs( coef[blkPos] ) * P.quantScale;
      const int64_t cap = (int64_t) INT32_MAX - ( (int64_t) 1 << ( qBits - 1 ) );
      const int32_t levelDouble = (int32_t)( tmpLevel < cap ? tmpLevel : cap );

      const uint32_t roundAbsLevel = (uint32_t) rq_min( (int) entropyCodingMaximum, (int)( (uint32_t)( levelDouble + ( (int32_t) 1 << ( qBits - 1 ) ) ) >> qBits ) );
      const uint32_t minAbsLevel = roundAbsLevel > 1 ? roundAbsLevel - 1 : 1;
      const uint32_t downAbsLevel = (uint32_t) rq_min( (int) entropyCodingMaximum, (int)( levelDouble >> qBits ) );
      const uint32_t upAbsLevel = (uint32_t) rq_min( (int) entropyCodingMaximum, (int)( downAbsLevel + 1 ) );

      uint32_t coeffLevels[3];
      int testedLevels = 0;
      coeffLevels[testedLevels++] = roundAbsLevel;
      if( minAbsLevel != roundAbsLevel ) coeffLevels[testedLevels++] = minAbsLevel;

      const int rightPixel = posX > 0 ? q[blkPos - 1] : 0;        // neighTS: the left and the upper neighbour (named as in the reference)
      const int belowPixel = posY > 0 ? q[blkPos - W] : 0;
      const int predPixel = rq_ts_mod_coeff( rightPixel, belowPixel, (int) upAbsLevel );
      if( upAbsLevel != roundAbsLevel && upAbsLevel != minAbsLevel && predPixel == 1 ) coeffLevels[testedLevels++] = upAbsLevel;

      const double dErr0 = (double) levelDouble;
      const double costCoeff0 = dErr0 * dErr0 * P.errorScale;

      // contexts from the two neighbours: significance and greater-1 count the non-zero ones, the sign context looks at their signs (ContextModelling.h:271-357)
      const int numPos = ( rightPixel != 0 ) + ( belowPixel != 0 );
      const int32_t* fbSig = R.sigBits[numPos];
      const int32_t* fbGt1 = R.lrg1Bits[numPos];
      int signCtx;
      if( ( rightPixel == 0 && belowPixel == 0 ) || ( rightPixel * belowPixel ) < 0 ) signCtx = 0;
      else if( rightPixel >= 0 && belowPixel >= 0 ) signCtx = 1;
      else signCtx = 2;
      const int32_t* fbSign = R.signBits[signCtx];
      const int sign = coef[blkPos] < 0 ? 1 : 0;
      const uint32_t goRiceParam = 1;
      const bool lastCoeff = scanPosInSB == 15 && noCoeffCoded == 0;

      // xGetCodedLevelTSPred, :1578-1661
      double costCoeff, costSig = 0.0;
      uint32_t cLevel = 0;
      int numUsedCtxBins = 0;
      {
        double currCostSig = 0;
        int numBestCtxBin = 0;
        bool done = false;
        if( !lastCoeff && coeffLevels[0] < 3 )
        {
          if( remRegBins >= 4 ) costSig = P.lambda * (double) fbSig[0];
          else                  costSig = P.lambda * (double)( 1 << RQ_SCALE_BITS );
          costCoeff = costCoeff0 + costSig;
          if( remRegBins >= 4 ) numUsedCtxBins++;
          if( coeffLevels[0] == 0 ) done = true;
        }
        else costCoeff = 1.7e+308;                                // MAX_DOUBLE (CommonDef.h)
        if( !done )
        {
          if( !lastCoeff )
          {
            if( remRegBins >= 4 ) currCostSig = P.lambda * (double) fbSig[1];
            else                  currCostSig = P.lambda * (double)( 1 << RQ_SCALE_BITS );
            if( coeffLevels[0] >= 3 && remRegBins >= 4 ) numUsedCtxBins++;
          }
          for( int errorInd = 1; errorInd <= testedLevels; errorInd++ )
          {
            const int absLevel = (int) coeffLevels[errorInd - 1];
            const double dErr = (double)( levelDouble - ( (int32_t) absLevel << qBits ) );
            const double levelError = dErr * dErr * P.errorScale;
            int modAbsLevel = absLevel;
            if( remRegBins >= 4 ) modAbsLevel = rq_ts_mod_coeff( rightPixel, belowPixel, absLevel );
            int numCtxBins = 0;
            double dCurrCost = levelError + P.lambda * (double) rq_ts_level_rate( R, (uint32_t) modAbsLevel, remRegBins, fbSign, fbGt1, numCtxBins, sign, goRiceParam );
            if( remRegBins >= 4 ) dCurrCost += currCostSig;
            if( dCurrCost < costCoeff ) { cLevel = (uint32_t) absLevel; costCoeff = dCurrCost; costSig = currCostSig; numBestCtxBin = numCtxBins; }
          }
          numUsedCtxBins += numBestCtxBin;
        }
      }

      remRegBins -= numUsedCtxBins;
      iNumSbbCtxBins += numUsedCtxBins;
      if( cLevel > 0 ) noCoeffCoded++;
      const int level = (int) cLevel;
      q[blkPos] = (int16_t)( ( level != 0 && coef[blkPos] < 0 ) ? -level : level );
      baseCost   += costCoeff;
      d64SigCost += costSig;
      if( q[blkPos] )
      {
        sigGroupFlags |= cgBit;
        d64CodedLevelandDist += costCoeff - costSig;
        d64UncodedDist       += costCoeff0;
      }
    }

    if( !( sigGroupFlags & cgBit ) )                              // :1271-1277
    {
      baseCost += P.lambda * (double) fbSigGroup[0] - d64SigCost;
      remRegBins += iNumSbbCtxBins;
    }
    else if( sbId != sbNum - 1 || anySigCG )                      // :1278-1322
    {
      double costZeroSB = baseCost;
      baseCost   += P.lambda * (double) fbSigGroup[1];
      costZeroSB += P.lambda * (double) fbSigGroup[0];
      costZeroSB += d64UncodedDist;
      costZeroSB -= d64CodedLevelandDist;
      costZeroSB -= d64SigCost;
      if( costZeroSB < baseCost )
      {
        sigGroupFlags &= ~cgBit;
        baseCost = costZeroSB;
        remRegBins += iNumSbbCtxBins;
        for( int p = 0; p <= 15; p++ )
        {
          const int raster = scan[( sbId << 4 ) + p];
          q[( ( raster >> lrw ) << lw ) + ( raster & ( regionW - 1 ) )] = 0;
        }
      }
      else anySigCG = true;
    }
  }

  for( int i = 0; i < W * H; i++ ) absSum += rq_abs( q[i] );     // :1325-1335 (every position is inside the scan: transform skip exists up to 32 x 32)
  *absSumOut = absSum;
}


// ------------------------------------------------------------------------------------------------------------------------------------------------------------------
// BDPCM: QuantRDOQ::forwardRDPCM (CommonLib/QuantRDOQ.cpp:1338-1562), the quantiser of a transform-skipped TU whose CU carries a block-DPCM direction (1 horizontal, 2 vertical).
// The routine of rq_ts_quant_tu with three differences: what is quantised is the residual minus the RECONSTRUCTED left / upper neighbour (xDequantSample :1564-1576 of the level
// just chosen plus its own prediction, kept in fullCoeff), the contexts take their BDPCM variants (greater-1: numPos 3; sign: + 3; no neighbour-based level mapping), and only
// the rounded level and the one below are tried.  fullCoeff: w * h int32 of scratch per TU.  One quirk is kept on purpose: when a group is zeroed out, the member refreshes
// m_fullCoeff at index scanPos instead of blkPos (:1539) -- the reconstruction other positions predict from is the one the member has.
struct RqBdpcmPar
{
  int32_t dirMode;                   // tu.cu->bdpcmM[chType]: 1 horizontal, 2 vertical
  int32_t dqScale;                   // g_invQuantScales[0][ qp.rem( true ) ], :1383
  int32_t dqRightShift;              // IQUANT_SHIFT - qp.per( true ), :1382
  int32_t pad;
};

VVB_HD int32_t rq_dequant_sample( int level, const RqBdpcmPar& B )          // xDequantSample, :1564-1576
{
  if( B.dqRightShift > 0 )
  {
    const int32_t qAdd = (int32_t) 1 << ( B.dqRightShift - 1 );
    return (int32_t)( ( (int32_t) level * B.dqScale + qAdd ) >> B.dqRightShift );
  }
  return (int32_t)( ( (int32_t) level * B.dqScale ) * ( 1 << -B.dqRightShift ) );
}

VVB_HD void rq_bdpcm_quant_tu( const RqTsPar& P, const RqBdpcmPar& B, const RqTsRates& R, const int32_t* scan, const int32_t* coef, int16_t* q, int32_t* fullCoeff, int32_t* absSumOut )
{
  const int W = P.width, H = P.height, lw = P.log2W;
  const int regionW = rq_min( 32, W );
  const int lrw = ( regionW == 32 ? 5 : regionW == 16 ? 4 : regionW == 8 ? 3 : 2 );
  const int qBits = P.qBits;
  const int widthInGroups = W >> 2;
  const int sbNum = ( W * H ) >> 4;
  const uint32_t entropyCodingMaximum = ( 1u << 15 ) - 1;
  const int dirMode = B.dirMode;
  uint64_t sigGroupFlags = 0;
  bool anySigCG = false;
  int remRegBins = P.maxCtxBins;
  int absSum = 0;

  for( int i = 0; i < W * H; i++ ) { q[i] = 0; fullCoeff[i] = 0; }       // :1368-1370

  for( int sbId = 0; sbId < sbNum; sbId++ )
  {
    const int cgRaster = scan[sbId << 4], cgX = ( cgRaster & ( regionW - 1 ) ) >> 2, cgY = ( cgRaster >> lrw ) >> 2;
    const int subSetPos = cgY * widthInGroups + cgX;
    const uint64_t cgBit = (uint64_t) 1 << subSetPos;
    const int sigLeft  = cgX > 0 ? (int)( ( sigGroupFlags >> ( subSetPos - 1 ) ) & 1 ) : 0;
    const int sigAbove = cgY > 0 ? (int)( ( sigGroupFlags >> ( subSetPos - widthInGroups ) ) & 1 ) : 0;
    const int32_t* fbSigGroup = R.sigGroupBits[sigLeft + sigAbove];

    int noCoeffCoded = 0;
    double baseCost = 0.0;
    double d64CodedLevelandDist = 0.0, d64UncodedDist = 0.0, d64SigCost = 0.0;
    int iNumSbbCtxBins = 0;

    for( int scanPosInSB = 0; scanPosInSB <= 15; scanPosInSB++ )
    {
      const int scanPos = ( sbId << 4 ) + scanPosInSB;
      const int raster = scan[scanPos], posX = raster & ( regionW - 1 ), posY = raster >> lrw;
      const int blkPos = ( posY << lw ) + posX;
      const int posS = ( 1 == dirMode ) ? posX : posY;
      const int posNb = ( 1 == dirMode ) ? ( posX - 1 ) + posY * W : posX + ( posY - 1 ) * W;
      const int32_t predCoeff = ( 0 != posS ) ? fullCoeff[posNb] : 0;

      const int64_t tmpLevel = (int64_t) rq_abs( coef[blkPos] - predCoeff ) * P.quantScale;
      const int64_t cap = (int64_t) INT32_MAX - ( (int64_t) 1 << ( qBits - 1 ) );
      const int32_t levelDouble = (int32_t)( tmpLevel < cap ? tmpLevel : cap );
      const uint32_t roundAbsLevel = (uint32_t) rq_min( (int) entropyCodingMaximum, (int)( (uint32_t)( levelDouble + ( (int32_t) 1 << ( qBits - 1 ) ) ) >> qBits ) );
      const uint32_t minAbsLevel = roundAbsLevel > 1 ? roundAbsLevel - 1 : 1;
      uint32_t coeffLevels[3];
      int testedLevels = 0;
      coeffLevels[testedLevels++] = roundAbsLevel;
      if( minAbsLevel != roundAbsLevel ) coeffLevels[testedLevels++] = minAbsLevel;

      const double dErr0 = (double) levelDouble;
      const double costCoeff0 = dErr0 * dErr0 * P.errorScale;

      const int rightPixel = posX > 0 ? q[blkPos - 1] : 0;
      const int belowPixel = posY > 0 ? q[blkPos - W] : 0;
      const int numPos = ( rightPixel != 0 ) + ( belowPixel != 0 );
      const int32_t* fbSig = R.sigBits[numPos];                  // sigCtxIdAbsTS has no BDPCM variant
      const int32_t* fbGt1 = R.lrg1Bits[3];                      // lrg1CtxIdAbsTS( ., ., bdpcm ): numPos = 3
      int signCtx;
      if( ( rightPixel == 0 && belowPixel == 0 ) || ( rightPixel * belowPixel ) < 0 ) signCtx = 0;
      else if( rightPixel >= 0 && belowPixel >= 0 ) signCtx = 1;
      else signCtx = 2;
      const int32_t* fbSign = R.signBits[signCtx + 3];           // signCtxIdAbsTS( ., ., bdpcm ): + 3
      const int sign = coef[blkPos] - predCoeff < 0 ? 1 : 0;
      const uint32_t goRiceParam = 1;
      const bool lastCoeff = scanPosInSB == 15 && noCoeffCoded == 0;

      double costCoeff, costSig = 0.0;
      uint32_t cLevel = 0;
      int numUsedCtxBins = 0;
      {
        double currCostSig = 0;
        int numBestCtxBin = 0;
        bool done = false;
        if( !lastCoeff && coeffLevels[0] < 3 )
        {
          if( remRegBins >= 4 ) costSig = P.lambda * (double) fbSig[0];
          else                  costSig = P.lambda * (double)( 1 << RQ_SCALE_BITS );
          costCoeff = costCoeff0 + costSig;
          if( remRegBins >= 4 ) numUsedCtxBins++;
          if( coeffLevels[0] == 0 ) done = true;
        }
        else costCoeff = 1.7e+308;
        if( !done )
        {
          if( !lastCoeff )
          {
            if( remRegBins >= 4 ) currCostSig = P.lambda * (double) fbSig[1];
            else                  currCostSig = P.lambda * (double)( 1 << RQ_SCALE_BITS );
            if( coeffLevels[0] >= 3 && remRegBins >= 4 ) numUsedCtxBins++;
          }
          for( int errorInd = 1; errorInd <= testedLevels; errorInd++ )
          {
            const int absLevel = (int) coeffLevels[errorInd - 1];
            const double dErr = (double)( levelDouble - ( (int32_t) absLevel << qBits ) );
            const double levelError = dErr * dErr * P.errorScale;
            int numCtxBins = 0;                                   // deriveModCoeff( ., ., absLevel, bdpcm != 0 ) leaves the level as it is
            double dCurrCost = levelError + P.lambda * (double) rq_ts_level_rate( R, (uint32_t) absLevel, remRegBins, fbSign, fbGt1, numCtxBins, sign, goRiceParam );
            if( remRegBins >= 4 ) dCurrCost += currCostSig;
            if( dCurrCost < costCoeff ) { cLevel = (uint32_t) absLevel; costCoeff = dCurrCost; costSig = currCostSig; numBestCtxBin = numCtxBins; }
          }
          numUsedCtxBins += numBestCtxBin;
        }
      }

      remRegBins -= numUsedCtxBins;
      iNumSbbCtxBins += numUsedCtxBins;
      if( cLevel > 0 ) noCoeffCoded++;
      q[blkPos] = (int16_t)( sign ? -(int) cLevel : (int) cLevel );
      fullCoeff[blkPos] = rq_dequant_sample( q[blkPos], B ) + predCoeff;          // :1491-1492
      baseCost   += costCoeff;
      d64SigCost += costSig;
      if( q[blkPos] )
      {
        sigGroupFlags |= cgBit;
        d64CodedLevelandDist += costCoeff - costSig;
        d64UncodedDist       += costCoeff0;
      }
    }

    if( !( sigGroupFlags & cgBit ) )
    {
      baseCost += P.lambda * (double) fbSigGroup[0] - d64SigCost;
      remRegBins += iNumSbbCtxBins;
    }
    else if( sbId != sbNum - 1 || anySigCG )
    {
      double costZeroSB = baseCost;
      baseCost   += P.lambda * (double) fbSigGroup[1];
      costZeroSB += P.lambda * (double) fbSigGroup[0];
      costZeroSB += d64UncodedDist;
      costZeroSB -= d64CodedLevelandDist;
      costZeroSB -= d64SigCost;
      if( costZeroSB < baseCost )
      {
        sigGroupFlags &= ~cgBit;
        baseCost = costZeroSB;
        remRegBins += iNumSbbCtxBins;
        for( int p = 0; p <= 15; p++ )
        {
          const int scanPos = ( sbId << 4 ) + p;
          const int raster = scan[scanPos], posX = raster & ( regionW - 1 ), posY = raster >> lrw;
          const int blkPos = ( posY << lw ) + posX;
          const int posS = ( 1 == dirMode ) ? posX : posY;
          const int posNb = ( 1 == dirMode ) ? ( posX - 1 ) + posY * W : posX + ( posY - 1 ) * W;
          fullCoeff[scanPos] = ( 0 != posS ) ? fullCoeff[posNb] : 0;              // the member indexes by scanPos here (:1539)
          q[blkPos] = 0;
        }
      }
      else anySigCG = true;
    }
  }

  for( int i = 0; i < W * H; i++ ) absSum += rq_abs( q[i] );
  *absSumOut = absSum;
}

} // namespace vvbrq
