// integration/TrQuantB200.h -- reference-side binding of libvvenc_b200.so for the transform / quantisation seam of CommonLib/TrQuant.cpp.
//
//   xTQuantB200          <->  the xT + xQuant pair inside TrQuant::transformNxN (TrQuant.cpp:709-733 -> xT :481-564, Quant::quant Quant.cpp:735-833)
//   invTransformNxNB200  <->  TrQuant::invTransformNxN (TrQuant.cpp:318-348 = Quant::dequant Quant.cpp:520-609 + xIT :567-660)
//   xQuantDQB200         <->  DepQuant::xQuantDQ (DepQuant.cpp:1129-1264), the trellis DepQuant::quant runs for non-skip TUs of a slice with depQuantEnabled
//   xRateDistOptQuantB200 <-> QuantRDOQ2::xRateDistOptQuant (QuantRDOQ2.cpp:1283-1296 -> xRateDistOptQuantFast :475-1281), the fast RDOQ of m_RDOQ == 2
//   rateDistOptQuantTSB200 <-> QuantRDOQ::rateDistOptQuantTS (QuantRDOQ.cpp:1124-1336), the RDOQ of transform-skipped TUs (m_useRDOQTS)
//   forwardRDPCMB200       <-> QuantRDOQ::forwardRDPCM (QuantRDOQ.cpp:1338-1562), the quantiser of BDPCM TUs; their inverse path inside invTransformNxNB200
//
// for the TUs the library covers: luma and chroma components, DCT-II / DST-VII / DCT-VIII (explicit MTS and the implicit / SBT choices xSetTrTypes makes), transform skip,
// LFNST (luma, and the chroma TUs of a separate tree), joint Cb-Cr TUs, no scaling lists / BDPCM / ACT, plain quantiser incl. its sign-bit hiding (the RDOQ variants are separate calls below and use the coefficients
// this call leaves in the temp buffer; dependent quantisation: xQuantDQB200 below).
// vvb_tu_par is derived from the TransformUnit exactly as the members derive their parameters (xSetTrTypes, QpParam, slice type), so the call sites keep
// their arguments.  One TU per call here; the production shape batches the TU candidates of a CU (vvb_fwd_trquant with n > 1 or
// vvb_tu_roundtrip).  Include after RdCostB200.h and CommonLib/TrQuant.h; TrQuant::xSetTrTypes is private: inside the encoder these are member functions.
#pragma once
#include <vector>
#include "RdCostB200.h"

struct B200TuApi
{
  bool bound = false;
  decltype( &vvb_fwd_trquant )  fwdTrQuant = nullptr;
  decltype( &vvb_inv_trquant )  invTrQuant = nullptr;
  decltype( &vvb_dep_quant )    depQuant   = nullptr;
  decltype( &vvb_rdoq )         rdoq       = nullptr;
  decltype( &vvb_rdoq_ts )      rdoqTs     = nullptr;
  decltype( &vvb_rdoq_bdpcm )   rdoqBdpcm  = nullptr;
} ;
static B200TuApi g_b200t;

inline int b200LoadTu( const char* libPath )
{
  if( g_b200t.bound ) return 0;
  int rc = b200Load( libPath );
  if( rc ) return rc;
  void* h = g_b200.handle;
#define VVB_RESOLVE( member, name ) g_b200t.member = (decltype( g_b200t.member )) dlsym( h, #name ); if( !g_b200t.member ) { g_b200.error = "missing " #name; return -2; }
  VVB_RESOLVE( fwdTrQuant, vvb_fwd_trquant )  VVB_RESOLVE( invTrQuant, vvb_inv_trquant )  VVB_RESOLVE( depQuant, vvb_dep_quant )  VVB_RESOLVE( rdoq, vvb_rdoq )  VVB_RESOLVE( rdoqTs, vvb_rdoq_ts )  VVB_RESOLVE( rdoqBdpcm, vvb_rdoq_bdpcm )
#undef VVB_RESOLVE
  g_b200t.bound = true;
  return 0;
}

inline vvb_tu_par b200TuPar( TrQuant& tq, const TransformUnit& tu, const ComponentID compID, const QpParam& cQP, bool forward = true, bool bdpcmOk = false )
{
  const ChannelType chType = toChannelType( compID );
  if( tu.cu->bdpcmM[chType] && !bdpcmOk ) THROW( "BDPCM TUs: forwardRDPCMB200 / invTransformNxNB200" );
  if( tu.cs->sps->scalingListEnabled ) THROW( "scaling lists stay on the host" );
  // joint Cb-Cr TUs need nothing special here: the caller has already formed the joint residual (fwdTransformICT) and built cQP for the joint mode (QpParam, Quant.cpp:80-87);
  // neither transformNxN nor the quantisers look at tu.jointCbCr.  The adaptive colour transform changes QpParam only when the caller asks for it: left to the host.
  if( tu.cu->colorTransform ) THROW( "ACT residuals stay on the host" );
  const SPS& sps = *tu.cs->sps;
  const bool skip = tu.mtsIdx[compID] == MTS_SKIP;
  int trHor = DCT2, trVer = DCT2;
  if( !skip ) tq.xSetTrTypes( tu, compID, tu.blocks[compID].width, tu.blocks[compID].height, trHor, trVer );     // TrQuant.cpp:417-478
  vvb_tu_par par = {};
  par.w = tu.blocks[compID].width; par.h = tu.blocks[compID].height;
  par.tr_hor = trHor; par.tr_ver = trVer;                                                    // enum TransType: DCT2 0, DCT8 1, DST7 2 -- the ABI's numbering
  par.bit_depth = sps.bitDepths[chType];
  // cQP.Qp( false ) is QpParam's base QP -- for chroma already mapped through the chroma QP table and offsets (Quant.cpp:96-113); the library re-adds 6 * (bitDepth - 8)
  par.qp = cQP.Qp( false ) - sps.qpBDOffset[chType];
  if( sps.qpBDOffset[chType] != 6 * ( par.bit_depth - 8 ) ) THROW( "unexpected qpBDOffset" );
  par.is_chroma = isChroma( compID ) ? 1 : 0;                                               // Quant::xNeedRDOQ rounds with 256 for chroma (Quant.cpp:877)
  par.transform_skip = skip ? 1 : 0;                                                        // xTransformSkip / xITransformSkip, QP floor 4 + 6 * internalMinusInputBitDepth (Quant.cpp:117-124)
  par.input_bit_depth_delta = sps.internalMinusInputBitDepth[chType];
  par.is_irap = tu.cs->slice->isIRAP() ? 1 : 0;                                              // rounding offset 171 vs 85 (Quant.cpp:772)
  par.dep_quant = tu.cs->slice->depQuantEnabled ? 1 : 0;                                     // xNeedRDOQ's QP; invTransformNxNB200 then dequantises as DepQuant::dequant does
  if( tu.cu->lfnstIdx && tu.cs->sps->LFNST && !skip && ( isLuma( compID ) || CU::isSepTree( *tu.cu ) ) )         // TrQuant::xFwdLfnst (TrQuant.cpp:942-1048) / xInvLfnst (:838-940): kernel set and transposition from the intra mode
  {
    if( trHor != DCT2 || trVer != DCT2 ) THROW( "LFNST index on a TU the library does not cover" );
    const CodingUnit& cu = *tu.cu;                                                           // :846 / :950 look the CU up at the TU position: the CU that owns the TU
    uint32_t intraMode = CU::getFinalIntraMode( cu, chType );                                // chroma TUs of a separate tree: the chroma mode, or for the cross-component modes
    if( CU::isLMCMode( cu.intraDir[chType] ) ) intraMode = CU::getCoLocatedIntraLumaMode( cu );   // the mode of the co-located luma CU (:958-961)
    if( CU::isMIP( cu, chType ) ) intraMode = PLANAR_IDX;
    intraMode = tq.xGetLFNSTIntraMode( ( tu.cu->ispMode && isLuma( compID ) ) ? tu.cu->blocks[compID] : tu.blocks[compID], intraMode );
    par.lfnst_idx = tu.cu->lfnstIdx; par.lfnst_set = g_lfnstLut[intraMode]; par.lfnst_transpose = tq.xGetTransposeFlag( intraMode ) ? 1 : 0;
  }
  par.sign_hiding = tu.cs->slice->signDataHidingEnabled ? 1 : 0;                            // Quant::quant: CoeffCodingContext( ..., signDataHidingEnabled ), xSignBitHidingHDQ (Quant.cpp:748, 817-826)
  return par;
}

// xT( tu, compID, resiBuf, tempCoeff, w, h ) followed by xQuant( tu, compID, tempCoeff, uiAbsSum, cQP, ctx ) with the plain quantiser:
// tempCoeff receives the transform coefficients, tu.getCoeffs( compID ) the levels, tu.lastPos[compID] and uiAbsSum as Quant::quant sets them.
// needRdoq (nullable) receives Quant::xNeedRDOQ of the coefficients (Quant.cpp:835-891).
inline void xTQuantB200( TrQuant& tq, TransformUnit& tu, const ComponentID compID, const CPelBuf& resiBuf, CoeffBuf& tempCoeff, const QpParam& cQP, TCoeff& uiAbsSum,
                         bool* needRdoq = nullptr, bool bdpcmOk = false )
{
  // bdpcmOk: the caller quantises the coefficients itself (forwardRDPCMB200); the levels of this call carry no DPCM and are to be overwritten
  const vvb_tu_par par = b200TuPar( tq, tu, compID, cQP, true, bdpcmOk );
  const int w = par.w, h = par.h;
  std::vector<int16_t> resi( (size_t) w * h ), q( (size_t) w * h );
  std::vector<int32_t> coef( (size_t) w * h );
  for( int y = 0; y < h; y++ ) memcpy( &resi[(size_t) y * w], resiBuf.buf + (ptrdiff_t) y * resiBuf.stride, sizeof( int16_t ) * w );
  int32_t absSum = 0, lastPos = -1; uint8_t nr = 0;
  b200Check( g_b200t.fwdTrQuant( b200CtxOfThread(), &par, resi.data(), 1, coef.data(), q.data(), &absSum, &lastPos, &nr ) );
  if( tu.cu->lfnstIdx && !par.lfnst_idx && !par.transform_skip )
  {
    // xT zeroes everything outside the top-left 4x4 / 8x8 whenever the CU carries an LFNST index (TrQuant.cpp:499-511) -- also for the chroma TUs of a single-tree CU, to
    // which the LFNST itself does not apply (:948).  Zero-out only drops outputs of the separable transform, so it is applied here to the full transform the library
    // returned; the levels of this call are not used by such callers (they quantise the coefficients afterwards), they are cleared with the coefficients
    const int keep = ( ( w == 4 && h > 4 ) || ( w > 4 && h == 4 ) ) ? 4 : ( ( w >= 8 && h >= 8 ) ? 8 : 0 );
    if( keep )
      for( int y = 0; y < h; y++ )
        for( int x = 0; x < w; x++ )
          if( x >= keep || y >= keep ) { coef[(size_t) y * w + x] = 0; q[(size_t) y * w + x] = 0; }
  }
  for( int y = 0; y < h; y++ ) memcpy( tempCoeff.buf + (ptrdiff_t) y * tempCoeff.stride, &coef[(size_t) y * w], sizeof( TCoeff ) * w );
  CoeffSigBuf dst = tu.getCoeffs( compID );
  for( int y = 0; y < h; y++ ) memcpy( dst.buf + (ptrdiff_t) y * dst.stride, &q[(size_t) y * w], sizeof( TCoeffSig ) * w );
  tu.lastPos[compID] = lastPos;
  uiAbsSum = absSum;
  if( needRdoq ) *needRdoq = nr != 0;
}

// TrQuant::invTransformNxN( tu, compID, pResi, cQP ) for the same class of TUs: levels of tu.getCoeffs( compID ) -> residual
inline void invTransformNxNB200( TrQuant& tq, TransformUnit& tu, const ComponentID compID, PelBuf& pResi, const QpParam& cQP )
{
  const vvb_tu_par par = b200TuPar( tq, tu, compID, cQP, false, true );
  const int w = par.w, h = par.h;
  std::vector<int16_t> q( (size_t) w * h ), resi( (size_t) w * h );
  const CCoeffSigBuf src = tu.getCoeffs( compID );
  for( int y = 0; y < h; y++ ) memcpy( &q[(size_t) y * w], src.buf + (ptrdiff_t) y * src.stride, sizeof( TCoeffSig ) * w );
  if( const int dirMode = tu.cu->bdpcmM[toChannelType( compID )] )
  {
    // Quant::dequant undoes the DPCM on the levels before the dequantiser of skipped transforms (invResDPCM, Quant.cpp:298-340): clipped running sums along the direction
    if( !par.transform_skip ) THROW( "BDPCM TU without MTS_SKIP" );
    const int lo = -( 1 << 15 ), hi = ( 1 << 15 ) - 1;
    if( dirMode == 1 ) { for( int y = 0; y < h; y++ ) for( int x = 1; x < w; x++ ) q[(size_t) y * w + x] = (int16_t) Clip3( lo, hi, int( q[(size_t) y * w + x - 1] ) + int( q[(size_t) y * w + x] ) ); }
    else               { for( int y = 1; y < h; y++ ) for( int x = 0; x < w; x++ ) q[(size_t) y * w + x] = (int16_t) Clip3( lo, hi, int( q[(size_t) ( y - 1 ) * w + x] ) + int( q[(size_t) y * w + x] ) ); }
  }
  b200Check( g_b200t.invTrQuant( b200CtxOfThread(), &par, q.data(), 1, resi.data() ) );
  for( int y = 0; y < h; y++ ) memcpy( pResi.buf + (ptrdiff_t) y * pResi.stride, &resi[(size_t) y * w], sizeof( Pel ) * w );
}

// DepQuant::xQuantDQ( tu, srcCoeff, compID, cQP, lambda, ctx, absSum, false, nullptr ) with the trellis on the device.  What stays here is what depends on the
// encoder's entropy-coding state: RateEstimator::initCtx (the member DepQuant inherits) turns the CABAC contexts into the rate tables, which travel as
// vvb_dq_rates.  Inside the encoder this is a member of DepQuant (RateEstimator is a private base); `dq` is that object.  Scaling lists stay on the host.
#include "CommonLib/DepQuant.h"
inline void xQuantDQB200( DepQuant& dq, TrQuant& tq, TransformUnit& tu, const CCoeffBuf& srcCoeff, const ComponentID compID, const QpParam& cQP, const double lambda, const Ctx& ctx, TCoeff& absSum )
{
  vvb_tu_par par = b200TuPar( tq, tu, compID, cQP );
  const int w = par.w, h = par.h;
  const DQIntern::TUParameters& tuPars = *dq.m_scansRom->getTUPars( tu.blocks[compID], compID );
  DQIntern::RateEstimator& re = (DQIntern::RateEstimator&) dq;
  re.initCtx( tuPars, tu, compID, ctx.getFracBitsAcess() );                                   // DepQuant.cpp:1199 (done before the first-position test here: the tables are inputs of the call)
  vvb_dq_rates rates;
  // lastOffset( scanIdx ) = m_lastBitsX[x] + m_lastBitsY[y] (DepQuant.h:167-170): split along the first row / first column, the constant part goes to x
  const int rw = std::min( w, 32 ), rh = std::min( h, 32 );
  std::vector<int> scanOf( (size_t) w * h, -1 );
  for( unsigned i = 0; i < tuPars.m_numCoeff; i++ ) scanOf[tuPars.m_scanId2BlkPos[i].idx] = (int) i;
  memset( &rates, 0, sizeof( rates ) );
  const int32_t corner = re.lastOffset( scanOf[0] );
  for( int x = 0; x < rw; x++ ) rates.last_bits_x[x] = re.lastOffset( scanOf[x] );
  for( int y = 0; y < rh; y++ ) rates.last_bits_y[y] = re.lastOffset( scanOf[(size_t) y * w] ) - corner;
  for( int i = 0; i < 2; i++ ) for( int b = 0; b < 2; b++ ) rates.sig_sbb_bits[i][b] = re.sigSbbFracBits()[i].intBits[b];
  for( int st = 0; st < 3; st++ ) for( int i = 0; i < 12; i++ ) for( int b = 0; b < 2; b++ ) rates.sig_bits[st][i][b] = re.sigFlagBits( st + 1 )[i].intBits[b];   // sigFlagBits( k ) = set max( k - 1, 0 )
  for( int i = 0; i < 21; i++ ) for( int b = 0; b < 6; b++ ) rates.gtx_bits[i][b] = re.gtxFracBits()[i].bits[b];
  vvb_dq_par dp = {};
  dp.lambda = lambda; dp.dq_thr_val = dq.m_quant.m_DqThrVal;
  dp.zero_out = ( ( tu.mtsIdx[compID] > MTS_SKIP || ( tu.cs->sps->MTS && tu.cu->sbtInfo != 0 && h <= 32 && w <= 32 ) ) && compID == COMP_Y ) ? 1 : 0;     // DepQuant.cpp:1155
  par.lfnst_idx = ( tu.cu->lfnstIdx > 0 && tu.mtsIdx[compID] != MTS_SKIP ) ? tu.cu->lfnstIdx : 0;                                // :1164
  std::vector<int32_t> coef( (size_t) w * h );
  std::vector<int16_t> q( (size_t) w * h );
  for( int y = 0; y < h; y++ ) memcpy( &coef[(size_t) y * w], srcCoeff.buf + (ptrdiff_t) y * srcCoeff.stride, sizeof( TCoeff ) * w );
  int32_t sum = 0, lastPos = -1;
  b200Check( g_b200t.depQuant( b200CtxOfThread(), &par, &dp, &rates, coef.data(), nullptr, 1, q.data(), &sum, &lastPos ) );
  CoeffSigBuf dst = tu.getCoeffs( compID );
  for( int y = 0; y < h; y++ ) memcpy( dst.buf + (ptrdiff_t) y * dst.stride, &q[(size_t) y * w], sizeof( TCoeffSig ) * w );
  tu.lastPos[compID] = lastPos;
  absSum = sum;
}

// QuantRDOQ2::xRateDistOptQuant( tu, compID, pSrc, uiAbsSum, cQP, ctx, false ) with the level decisions on the device.  What stays here is what depends on the
// encoder's entropy-coding state: the fractional bits of the contexts the routine reads travel as vvb_rdoq_rates.  The last-position table is the member's own
// (xInitLastPosBitsTab, run under the member's condition so that a Cr TU after a coded Cb TU sees the table of the Cb call, QuantRDOQ2.cpp:490); the coded-block-flag
// context is resolved as :1185-1226 resolve it.  Inside the encoder this is a member of QuantRDOQ2; `rq` is that object (the DepQuant instance TrQuant owns).
// Transform-skipped TUs: rateDistOptQuantTSB200 below; BDPCM and scaling lists stay on the host.
#include "CommonLib/QuantRDOQ2.h"
inline void xRateDistOptQuantB200( QuantRDOQ2& rq, TrQuant& tq, TransformUnit& tu, const ComponentID compID, const CCoeffBuf& pSrc, TCoeff& uiAbsSum, const QpParam& cQP, const Ctx& ctx )
{
  if( tu.mtsIdx[compID] == MTS_SKIP ) THROW( "transform-skipped TUs go through rateDistOptQuantTSB200" );
  vvb_tu_par par = b200TuPar( tq, tu, compID, cQP );
  const int w = par.w, h = par.h;
  const ChannelType ch = toChannelType( compID );
  const FracBitsAccess& fb = ctx.getFracBitsAcess();
  const bool sbh = tu.cs->slice->signDataHidingEnabled;
  if( compID != COMP_Cr || !tu.cbf[COMP_Cb] )                                                   // :490
  {
    CoeffCodingContext cctx( tu, compID, sbh, false, rq.m_tplBuf );
    rq.xInitLastPosBitsTab( cctx, w, h, ch, fb );
  }
  vvb_rdoq_rates rates;
  memset( &rates, 0, sizeof( rates ) );
  for( int i = 0; i < (int) Ctx::SigFlag[ch].size() && i < 12; i++ )     for( int b = 0; b < 2; b++ ) rates.sig_bits[i][b] = fb.getFracBitsArray( Ctx::SigFlag[ch]( i ) ).intBits[b];
  for( int i = 0; i < (int) Ctx::ParFlag[ch].size() && i < 21; i++ )     for( int b = 0; b < 2; b++ ) rates.par_bits[i][b] = fb.getFracBitsArray( Ctx::ParFlag[ch]( i ) ).intBits[b];
  for( int i = 0; i < (int) Ctx::GtxFlag[ch + 2].size() && i < 21; i++ ) for( int b = 0; b < 2; b++ ) rates.gt1_bits[i][b] = fb.getFracBitsArray( Ctx::GtxFlag[ch + 2]( i ) ).intBits[b];
  for( int i = 0; i < (int) Ctx::GtxFlag[ch].size() && i < 21; i++ )     for( int b = 0; b < 2; b++ ) rates.gt2_bits[i][b] = fb.getFracBitsArray( Ctx::GtxFlag[ch]( i ) ).intBits[b];
  for( int i = 0; i < 2; i++ ) for( int b = 0; b < 2; b++ ) rates.sig_group_bits[i][b] = fb.getFracBitsArray( Ctx::SigCoeffGroup[ch]( i ) ).intBits[b];
  for( int i = 0; i < LAST_SIGNIFICANT_GROUPS; i++ ) { rates.last_bits_x[i] = rq.QuantRDOQ2::m_lastBitsX[ch][i]; rates.last_bits_y[i] = rq.QuantRDOQ2::m_lastBitsY[ch][i]; }
  if( !CU::isIntra( *tu.cu ) && isLuma( compID ) )                                              // :1185-1190
  {
    const BinFracBits f = fb.getFracBitsArray( Ctx::QtRootCbf() );
    rates.cbf_bits[0] = f.intBits[0]; rates.cbf_bits[1] = f.intBits[1];
  }
  else                                                                                          // :1191-1226
  {
    bool previousCbf = tu.cbf[COMP_Cb], lastCbfIsInferred = false;
    const bool useIntraSubPartitions = tu.cu->ispMode && isLuma( compID );
    if( useIntraSubPartitions )
    {
      bool rootCbfSoFar = false;
      const bool isLastSubPartition = CU::isISPLast( *tu.cu, tu.Y(), compID );
      const uint32_t nTus = tu.cu->ispMode == HOR_INTRA_SUBPARTITIONS ? tu.cu->lheight() >> Log2( tu.lheight() ) : tu.cu->lwidth() >> Log2( tu.lwidth() );
      if( isLastSubPartition )
      {
        TransformUnit* tuPointer = tu.cu->firstTU;
        for( int tuIdx = 0; tuIdx < (int) nTus - 1; tuIdx++ ) { rootCbfSoFar |= TU::getCbfAtDepth( *tuPointer, COMP_Y, tu.depth ); tuPointer = tuPointer->next; }
        if( !rootCbfSoFar ) lastCbfIsInferred = true;
      }
      if( !lastCbfIsInferred ) previousCbf = TU::getPrevTuCbfAtDepth( tu, compID, tu.depth );
    }
    if( !lastCbfIsInferred )
    {
      const BinFracBits f = fb.getFracBitsArray( Ctx::QtCbf[compID]( DeriveCtx::CtxQtCbf( tu.blocks[compID].compID, previousCbf, useIntraSubPartitions ) ) );
      rates.cbf_bits[0] = f.intBits[0]; rates.cbf_bits[1] = f.intBits[1];
    }
  }
  vvb_rdoq_par rp = {};
  rp.lambda = rq.m_dLambda; rp.thr_val = rq.m_thrVal;
  rp.sbt_zero_out = ( tu.cs->sps->MTS && tu.cu->sbtInfo != 0 && w <= 32 && h <= 32 && compID == COMP_Y ) ? 1 : 0;     // TransformUnit::getTbAreaAfterCoefZeroOut, Unit.cpp:580
  par.lfnst_idx = tu.cu->lfnstIdx;                                                              // the routine reads the CU's index for every component (:552-559); the set / transposition fields are not used
  par.sign_hiding = sbh ? 1 : 0;
  std::vector<int32_t> coef( (size_t) w * h );
  std::vector<int16_t> q( (size_t) w * h );
  for( int y = 0; y < h; y++ ) memcpy( &coef[(size_t) y * w], pSrc.buf + (ptrdiff_t) y * pSrc.stride, sizeof( TCoeff ) * w );
  int32_t sum = 0, lastPos = -1;
  b200Check( g_b200t.rdoq( b200CtxOfThread(), &par, &rp, &rates, coef.data(), nullptr, 1, q.data(), &sum, &lastPos ) );
  CoeffSigBuf dst = tu.getCoeffs( compID );
  for( int y = 0; y < h; y++ ) memcpy( dst.buf + (ptrdiff_t) y * dst.stride, &q[(size_t) y * w], sizeof( TCoeffSig ) * w );
  if( lastPos >= 0 ) tu.lastPos[compID] = lastPos;                                              // the member writes tu.lastPos only when it codes something (:1258)
  uiAbsSum = sum;
}

// QuantRDOQ::rateDistOptQuantTS( tu, compID, coeffs, absSum, qp, ctx ) with the level decisions on the device: the fractional bits of the six transform-skip context sets
// travel as vvb_rdoq_ts_rates; the levels come back signed, tu.lastPos is left alone as the member leaves it.  BDPCM (forwardRDPCM) stays on the host.
inline void rateDistOptQuantTSB200( QuantRDOQ& rq, TrQuant& tq, TransformUnit& tu, const ComponentID compID, const CCoeffBuf& coeffs, TCoeff& absSum, const QpParam& qp, const Ctx& ctx )
{
  if( tu.mtsIdx[compID] != MTS_SKIP ) THROW( "rateDistOptQuantTSB200 is for transform-skipped TUs" );
  const vvb_tu_par par = b200TuPar( tq, tu, compID, qp );                                       // THROWs for BDPCM; transform_skip and input_bit_depth_delta are set there
  const int w = par.w, h = par.h;
  const FracBitsAccess& fb = ctx.getFracBitsAcess();
  vvb_rdoq_ts_rates rates;
  for( int i = 0; i < 3; i++ ) for( int b = 0; b < 2; b++ ) rates.sig_bits[i][b] = fb.getFracBitsArray( Ctx::TsSigFlag( i ) ).intBits[b];
  for( int b = 0; b < 2; b++ ) rates.par_bits[b] = fb.getFracBitsArray( Ctx::TsParFlag( 0 ) ).intBits[b];
  for( int i = 0; i < 5; i++ ) for( int b = 0; b < 2; b++ ) rates.gtx_bits[i][b] = fb.getFracBitsArray( Ctx::TsGtxFlag( i ) ).intBits[b];
  for( int i = 0; i < 4; i++ ) for( int b = 0; b < 2; b++ ) rates.lrg1_bits[i][b] = fb.getFracBitsArray( Ctx::TsLrg1Flag( i ) ).intBits[b];
  for( int i = 0; i < 6; i++ ) for( int b = 0; b < 2; b++ ) rates.sign_bits[i][b] = fb.getFracBitsArray( Ctx::TsResidualSign( i ) ).intBits[b];
  for( int i = 0; i < 3; i++ ) for( int b = 0; b < 2; b++ ) rates.sig_group_bits[i][b] = fb.getFracBitsArray( Ctx::TsSigCoeffGroup( i ) ).intBits[b];
  std::vector<int32_t> coef( (size_t) w * h );
  std::vector<int16_t> q( (size_t) w * h );
  for( int y = 0; y < h; y++ ) memcpy( &coef[(size_t) y * w], coeffs.buf + (ptrdiff_t) y * coeffs.stride, sizeof( TCoeff ) * w );
  int32_t sum = 0;
  b200Check( g_b200t.rdoqTs( b200CtxOfThread(), &par, rq.m_dLambda, &rates, coef.data(), nullptr, 1, q.data(), &sum ) );
  CoeffSigBuf dst = tu.getCoeffs( compID );
  for( int y = 0; y < h; y++ ) memcpy( dst.buf + (ptrdiff_t) y * dst.stride, &q[(size_t) y * w], sizeof( TCoeffSig ) * w );
  absSum += sum;                                                                                // the member accumulates into the caller's sum (:1334)
}

// QuantRDOQ::forwardRDPCM( tu, compID, coeffs, absSum, qp, ctx ) for a TU whose CU carries a BDPCM direction: the same rate tables as rateDistOptQuantTSB200, the direction
// travels as an argument, the reconstruction chain runs on the device.  coeffs must be compact (stride == width), as TrQuant's temp buffer is.
inline void forwardRDPCMB200( QuantRDOQ& rq, TrQuant& tq, TransformUnit& tu, const ComponentID compID, const CCoeffBuf& coeffs, TCoeff& absSum, const QpParam& qp, const Ctx& ctx )
{
  const int dirMode = tu.cu->bdpcmM[toChannelType( compID )];
  if( !dirMode || tu.mtsIdx[compID] != MTS_SKIP ) THROW( "forwardRDPCMB200 is for BDPCM TUs (mtsIdx == MTS_SKIP)" );
  const vvb_tu_par par = b200TuPar( tq, tu, compID, qp, true, true );
  const int w = par.w, h = par.h;
  const FracBitsAccess& fb = ctx.getFracBitsAcess();
  vvb_rdoq_ts_rates rates;
  for( int i = 0; i < 3; i++ ) for( int b = 0; b < 2; b++ ) rates.sig_bits[i][b] = fb.getFracBitsArray( Ctx::TsSigFlag( i ) ).intBits[b];
  for( int b = 0; b < 2; b++ ) rates.par_bits[b] = fb.getFracBitsArray( Ctx::TsParFlag( 0 ) ).intBits[b];
  for( int i = 0; i < 5; i++ ) for( int b = 0; b < 2; b++ ) rates.gtx_bits[i][b] = fb.getFracBitsArray( Ctx::TsGtxFlag( i ) ).intBits[b];
  for( int i = 0; i < 4; i++ ) for( int b = 0; b < 2; b++ ) rates.lrg1_bits[i][b] = fb.getFracBitsArray( Ctx::TsLrg1Flag( i ) ).intBits[b];
  for( int i = 0; i < 6; i++ ) for( int b = 0; b < 2; b++ ) rates.sign_bits[i][b] = fb.getFracBitsArray( Ctx::TsResidualSign( i ) ).intBits[b];
  for( int i = 0; i < 3; i++ ) for( int b = 0; b < 2; b++ ) rates.sig_group_bits[i][b] = fb.getFracBitsArray( Ctx::TsSigCoeffGroup( i ) ).intBits[b];
  std::vector<int32_t> coef( (size_t) w * h );
  std::vector<int16_t> q( (size_t) w * h );
  for( int y = 0; y < h; y++ ) memcpy( &coef[(size_t) y * w], coeffs.buf + (ptrdiff_t) y * coeffs.stride, sizeof( TCoeff ) * w );
  int32_t sum = 0;
  b200Check( g_b200t.rdoqBdpcm( b200CtxOfThread(), &par, rq.m_dLambda, dirMode, &rates, coef.data(), nullptr, 1, q.data(), &sum ) );
  CoeffSigBuf dst = tu.getCoeffs( compID );
  for( int y = 0; y < h; y++ ) memcpy( dst.buf + (ptrdiff_t) y * dst.stride, &q[(size_t) y * w], sizeof( TCoeffSig ) * w );
  absSum += sum;
}
