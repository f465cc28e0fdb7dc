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

} // namespace vvbrq
