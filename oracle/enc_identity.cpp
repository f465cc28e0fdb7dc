/*
 * oracle/enc_identity.cpp -- TEST INFRASTRUCTURE, not product code.
 *
 * The reference's own pass criterion for a SIMD back end is whole-encoder bitstream identity (`--SIMD=SCALAR` vs default must give identical .vvc files,
 * cmake/modules/vvencTests.cmake:52-53).  This program applies the same criterion to the B200 back end: it drives the UNMODIFIED reference encoder
 * (oracle/_ref/libvvenc_ref.a, compiled in place from /root/reference by oracle/Makefile.ref) through its public C API (vvenc_encoder_create / open / encode)
 * twice-compatible: once as built (AVX2 tables), once with `installB200()` of integration/RdCostB200.h and integration/AffineGradientB200.h applied to every
 * RdCost / AffineGradientSearch instance the encoder creates, so that every xGetSAD / xGetSSE / xGetHADs / HAD_2SAD / mask-SAD / SADx5 / weighted-SSE call and
 * the affine gradient helpers run on the GPU through libvvenc_b200.so (one call per block: the verification shape, not the production shape).
 *
 * How the tables get installed without touching the reference sources: the library calls RdCost::initRdCostX86() at the end of RdCost::create()
 * (CommonLib/RdCost.cpp:137) and AffineGradientSearch::initAffineGradientSearchX86() in its constructor (AffineGradientSearch.cpp:75); both live in another
 * object file (x86/InitX86.cpp), so the linker's --wrap redirects those two calls to the functions below, which run the original selection and then -- when
 * asked to -- the 10-line `_initRdCostB200()` a maintainer would add.
 *
 * With a ninth argument `tu` the transform / quantisation seam is routed through the library as well: TrQuant::transformNxN and TrQuant::invTransformNxN are
 * called from other object files (IntraSearch.cpp, InterSearch.cpp, EncCu.cpp), so --wrap hands them to the functions below, which keep the members' structure
 * (TrQuant.cpp:688-736, 318-348) and replace
 *   xT / xTransformSkip (+ xFwdLfnst for luma)        by xTQuantB200      (integration/TrQuantB200.h -> vvb_fwd_trquant: the coefficients it leaves are what xQuant then reads),
 *   DepQuant::xQuantDQ (slices with depQuantEnabled)   by xQuantDQB200     (rate tables from the live CABAC contexts -> vvb_dep_quant: the 4-state trellis on the device),
 *   Quant::dequant / DepQuant::dequant + xIT / xITransformSkip  by invTransformNxNB200 (TUs without LFNST; the rest stays with the member),
 *   QuantRDOQ2::xRateDistOptQuant (Quant::m_RDOQ == 2: presets faster / fast, and slices without dependent quantisation)  by xRateDistOptQuantB200 (fractional bits of the
 *                                                       live CABAC contexts -> vvb_rdoq: level decisions, group zero-out, last position, sign-bit hiding on the device),
 * With `turdoq` the routing is the widest the bindings offer: LFNST also on the chroma TUs of a separate tree (kernel set from the chroma / co-located luma mode) and on ISP luma TUs,
 * the chroma TUs of single-tree LFNST CUs (xT's zero-out applied to the full transform), joint Cb-Cr TUs (the caller has formed the joint residual and QP).
 * Transform-skipped TUs (environment VVB_ENC_TS=1 makes the encoder try transform skip on every eligible TU): `turdoq` / `all` route QuantRDOQ::rateDistOptQuantTS through rateDistOptQuantTSB200.
 * BDPCM TUs (VVB_ENC_BDPCM=1 next to VVB_ENC_TS=1): forwardRDPCMB200 (-> vvb_rdoq_bdpcm) and, on the inverse side, invTransformNxNB200 (running sums on the host + the library's inverse
 * of skipped transforms).
 * With `all` additionally the block-matching errors of the MCTF pre-analysis: MCTF::initMCTF_X86 is wrapped like the RdCost one and the error pointers / m_calcVar answer from
 * the library per call (integration/MCTFB200.h: installB200( MCTF& )), under the unmodified MCTF::motionEstimationLuma control.
 * RDOQ of m_RDOQ == 1 on non-skipped TUs, TUs with a side below 4 (thin ISP partitions, 2-wide chroma) and everything the bindings THROW for (ACT, scaling lists) stay with the members.
 *
 * usage: enc_identity <in.yuv (8-bit 4:2:0)> <width> <height> <frames> <preset 0..4 (faster..slower)> <qp> <out.vvc> [path of libvvenc_b200.so -> B200 tables] [tu | turdoq | all]
 * prints one line: `ENC frames=<n> bytes=<n> fnv1a=<hex> dist_calls=<n> b200=<0|1> ... tu_fwd=<n> tu_dq=<n> tu_rdoq=<n> tu_rdoq_ts=<n> tu_bdpcm=<n> tu_inv=<n> tu_inv_lfnst=<n> tu_ref=<n>`
 */
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <cmath>
#include <vector>
#include <array>
#include <deque>
#include <list>
#include <map>
#include <set>
#include <string>
#include <sstream>
#include <iostream>
#include <fstream>
#include <memory>
#include <mutex>
#include <thread>
#include <atomic>
#include <algorithm>
#include <functional>
#include <condition_variable>
#include <chrono>
#include <bitset>
#include <limits>
#include <numeric>
#include <unordered_map>
#include <future>
#include <cassert>
#include <cstdarg>
#include <immintrin.h>

#define private public
#define protected public
#include "vvenc/vvenc.h"
#include "vvenc/vvencCfg.h"
#include "CommonLib/CommonDef.h"
#include "CommonLib/Unit.h"
#include "CommonLib/RdCost.h"
#include "CommonLib/AffineGradientSearch.h"
#include "CommonLib/UnitTools.h"
#include "CommonLib/Slice.h"
#include "CommonLib/Picture.h"
#include "CommonLib/CodingStructure.h"
#include "CommonLib/TrQuant.h"
#include "CommonLib/Quant.h"
#include "CommonLib/DepQuant.h"
#include "CommonLib/MCTF.h"
#include "EncoderLib/InterSearch.h"
#include "CommonLib/Rom.h"
#include "CommonLib/Contexts.h"
#undef private
#undef protected

using namespace vvenc;
#include "../integration/RdCostB200.h"
#include "../integration/AffineGradientB200.h"
#include "../integration/TrQuantB200.h"
#include "../integration/MCTFB200.h"

static bool               g_useB200 = false;
static std::atomic<long>  g_rdCostInstalls{ 0 }, g_affineInstalls{ 0 };

extern "C" void __real__ZN5vvenc6RdCost13initRdCostX86Ev( RdCost* );
extern "C" void __wrap__ZN5vvenc6RdCost13initRdCostX86Ev( RdCost* self )
{
  __real__ZN5vvenc6RdCost13initRdCostX86Ev( self );           // the x86 selection, as the library does it
  if( g_useB200 ) { installB200( *self ); g_rdCostInstalls++; }
}
extern "C" void __real__ZN5vvenc20AffineGradientSearch27initAffineGradientSearchX86Ev( AffineGradientSearch* );
extern "C" void __wrap__ZN5vvenc20AffineGradientSearch27initAffineGradientSearchX86Ev( AffineGradientSearch* self )
{
  __real__ZN5vvenc20AffineGradientSearch27initAffineGradientSearchX86Ev( self );
  if( g_useB200 ) { installB200( *self ); g_affineInstalls++; }
}

// MCTF: the constructor calls initMCTF_X86() (MCTF.cpp:572, defined in x86/InitX86.cpp:322); with the argument `all` the error pointers and m_calcVar are then pointed at
// the per-call trampolines of integration/MCTFB200.h, so the unmodified motion search of the pre-analysis runs on the library's block-matching errors
static bool               g_useMctf = false;
static std::atomic<long>  g_mctfInstalls{ 0 };
static std::atomic<unsigned long long> g_mctfCalls{ 0 };
static decltype( &vvb_mctf_error_batch ) g_realMctfErr = nullptr;
static int countingMctfErr( vvb_ctx* c, int po, int pr, const vvb_mctf_cand* cd, int n, int lowRes, int32_t* e ) { g_mctfCalls++; return g_realMctfErr( c, po, pr, cd, n, lowRes, e ); }
extern "C" void __real__ZN5vvenc4MCTF12initMCTF_X86Ev( MCTF* );
extern "C" void __wrap__ZN5vvenc4MCTF12initMCTF_X86Ev( MCTF* self )
{
  __real__ZN5vvenc4MCTF12initMCTF_X86Ev( self );
  if( g_useMctf ) { installB200( *self ); g_mctfInstalls++; }
}

// call counters: thunks between the binding's function pointers and the library (the binding itself stays as a maintainer would ship it)
static std::atomic<unsigned long long> g_distCalls{ 0 }, g_otherCalls{ 0 };
static decltype( &vvb_dist_block ) g_realDist = nullptr;
static uint64_t countingDist( vvb_ctx* c, int f, const int16_t* o, int so, const int16_t* u, int su, int w, int h, int bd, int ss, int* e ) { g_distCalls++; return g_realDist( c, f, o, so, u, su, w, h, bd, ss, e ); }
static decltype( &vvb_sad_x5_block ) g_realX5 = nullptr;
static int countingX5( vvb_ctx* c, const int16_t* o, int so, const int16_t* u, int su, int w, int h, int ss, int cc, uint64_t* out ) { g_otherCalls++; return g_realX5( c, o, so, u, su, w, h, ss, cc, out ); }

// ---- the transform / quantisation seam -------------------------------------------------------------------------------------------------------------------------
static bool g_useTu = false, g_useRdoq = false;      // g_useRdoq: argument `turdoq` -- the widest routing: also the fast RDOQ of m_RDOQ == 2, LFNST on the chroma TUs of a
                                                      // separate tree and on ISP luma TUs, joint Cb-Cr TUs (`tu` keeps the narrower routing of the first hardware runs)
static std::atomic<unsigned long long> g_tuFwd{ 0 }, g_tuDq{ 0 }, g_tuRdoq{ 0 }, g_tuRdoqTs{ 0 }, g_tuBdpcm{ 0 }, g_tuInv{ 0 }, g_tuInvLfnst{ 0 }, g_tuRef{ 0 };

extern "C" void __real__ZN5vvenc7TrQuant12transformNxNERNS_13TransformUnitENS_11ComponentIDERKNS_7QpParamERiRKNS_3CtxEb( TrQuant*, TransformUnit&, ComponentID, const QpParam&, TCoeff&, const Ctx&, bool );
extern "C" void __wrap__ZN5vvenc7TrQuant12transformNxNERNS_13TransformUnitENS_11ComponentIDERKNS_7QpParamERiRKNS_3CtxEb( TrQuant* self, TransformUnit& tu, ComponentID compID, const QpParam& cQP,
                                                                                                                         TCoeff& uiAbsSum, const Ctx& ctx, bool loadTr )
{
  const ChannelType chType = toChannelType( compID );
  const bool lfnstHere = tu.cs->sps->LFNST && tu.cu->lfnstIdx != 0;
  // what stays with the member: BDPCM, LFNST on stored coefficients, empty TUs; with `tu` (not `turdoq`) also LFNST on chroma / ISP CUs and joint Cb-Cr TUs
  if( !g_useTu || tu.noResidual || ( tu.cu->bdpcmM[chType] && !g_useRdoq ) || ( lfnstHere && ( loadTr || ( !g_useRdoq && ( !isLuma( compID ) || tu.cu->ispMode ) ) ) ) || ( !g_useRdoq && isChroma( compID ) && tu.jointCbCr ) || tu.cs->sps->scalingListEnabled )
  {
    g_tuRef++;
    __real__ZN5vvenc7TrQuant12transformNxNERNS_13TransformUnitENS_11ComponentIDERKNS_7QpParamERiRKNS_3CtxEb( self, tu, compID, cQP, uiAbsSum, ctx, loadTr );
    return;
  }
  const CompArea& rect = tu.blocks[compID];
  const CPelBuf resiBuf = tu.cs->getResiBuf( rect );
  const bool bdpcm = tu.cu->bdpcmM[chType] != 0;
  if( bdpcm ) tu.mtsIdx[compID] = MTS_SKIP;                                            // TrQuant.cpp:703-706
  uiAbsSum = 0;
  CoeffBuf tempCoeff( loadTr ? self->m_mtsCoeffs[tu.mtsIdx[compID]] : self->m_plTempCoeff, rect );      // loadTr: checktransformsNxN has left the coefficients (TrQuant.cpp:709)
  if( !loadTr )
  try
  {
    TCoeff plainSum = 0;
    xTQuantB200( *self, tu, compID, resiBuf, tempCoeff, cQP, plainSum, nullptr, bdpcm );   // xT / xTransformSkip (+ xFwdLfnst): tempCoeff as the member leaves it
    g_tuFwd++;
  }
  catch( std::exception& )                                                             // a TU the binding does not cover (joint Cb-Cr, ...): the member
  {
    g_tuRef++;
    __real__ZN5vvenc7TrQuant12transformNxNERNS_13TransformUnitENS_11ComponentIDERKNS_7QpParamERiRKNS_3CtxEb( self, tu, compID, cQP, uiAbsSum, ctx, loadTr );
    return;
  }
  // xQuant (TrQuant.cpp:730 -> DepQuant::quant, DepQuant.cpp:1462-1490)
  DepQuant* dq = dynamic_cast<DepQuant*>( self->m_quant );
  const bool selectiveSkip = dq && tu.cs->picture->useSelectiveRdoq && !dq->xNeedRDOQ( tu, compID, tempCoeff, cQP );
  if( dq && !selectiveSkip && tu.cs->slice->depQuantEnabled && tu.mtsIdx[compID] != MTS_SKIP )
  {
    uiAbsSum = 0;
    xQuantDQB200( *dq, *self, tu, tempCoeff, compID, cQP, dq->m_dLambda, ctx, uiAbsSum );
    g_tuDq++;
  }
  else
  {
    uiAbsSum = 0;
    // QuantRDOQ2::quant (QuantRDOQ2.cpp:247-301): m_RDOQ == 2, not transform skipped, sides above 2 -> xRateDistOptQuant, unless the selective pre-check says "all zero"
    bool done = false;
    if( g_useRdoq && dq && !selectiveSkip && dq->m_RDOQ == 2 && tu.mtsIdx[compID] != MTS_SKIP && rect.width > 2 && rect.height > 2 && !dq->getScalingListEnabled() )
    {
      try { xRateDistOptQuantB200( *dq, *self, tu, compID, tempCoeff, uiAbsSum, cQP, ctx ); g_tuRdoq++; done = true; }
      catch( std::exception& ) { uiAbsSum = 0; }
    }
    // ... and for a transform-skipped TU without BDPCM rateDistOptQuantTS when Quant::m_useRDOQTS is set (:263, 275-285)
    if( !done && g_useRdoq && dq && !selectiveSkip && dq->m_useRDOQTS && tu.mtsIdx[compID] == MTS_SKIP && rect.width > 2 && rect.height > 2 )
    {
      try
      {
        if( bdpcm ) { forwardRDPCMB200( *dq, *self, tu, compID, tempCoeff, uiAbsSum, cQP, ctx ); g_tuBdpcm++; }
        else        { rateDistOptQuantTSB200( *dq, *self, tu, compID, tempCoeff, uiAbsSum, cQP, ctx ); g_tuRdoqTs++; }
        done = true;
      }
      catch( std::exception& ) { uiAbsSum = 0; }
    }
    if( !done ) self->xQuant( tu, compID, tempCoeff, uiAbsSum, cQP, ctx );
  }
  TU::setCbfAtDepth( tu, compID, tu.depth, uiAbsSum > 0 );
}

extern "C" void __real__ZN5vvenc7TrQuant15invTransformNxNERNS_13TransformUnitENS_11ComponentIDERNS_7AreaBufIsEERKNS_7QpParamE( TrQuant*, TransformUnit&, ComponentID, PelBuf&, const QpParam& );
extern "C" void __wrap__ZN5vvenc7TrQuant15invTransformNxNERNS_13TransformUnitENS_11ComponentIDERNS_7AreaBufIsEERKNS_7QpParamE( TrQuant* self, TransformUnit& tu, ComponentID compID, PelBuf& pResi, const QpParam& cQP )
{
  const ChannelType chType = toChannelType( compID );
  const bool lfnstHere = tu.cs->sps->LFNST && tu.cu->lfnstIdx != 0 && tu.mtsIdx[compID] != MTS_SKIP && ( CU::isSepTree( *tu.cu ) ? true : isLuma( compID ) );
  // plain and DepQuant dequantiser alike (vvb_tu_par.dep_quant); LFNST as in the forward wrapper
  if( g_useTu && ( g_useRdoq || !( ( lfnstHere && ( !isLuma( compID ) || tu.cu->ispMode ) ) || ( isChroma( compID ) && tu.jointCbCr ) || tu.cu->bdpcmM[chType] ) ) && !tu.cs->sps->scalingListEnabled )
  {
    try { invTransformNxNB200( *self, tu, compID, pResi, cQP ); g_tuInv++; if( lfnstHere ) g_tuInvLfnst++; return; }
    catch( std::exception& ) {}
  }
  g_tuRef++;
  __real__ZN5vvenc7TrQuant15invTransformNxNERNS_13TransformUnitENS_11ComponentIDERNS_7AreaBufIsEERKNS_7QpParamE( self, tu, compID, pResi, cQP );
}

static void quietLog( void*, int, const char*, va_list ) {}

int main( int argc, char** argv )
{
  if( argc < 8 ) { fprintf( stderr, "usage: %s in.yuv w h frames preset qp out.vvc [libvvenc_b200.so]\n", argv[0] ); return 2; }
  const char* inPath = argv[1];
  const int w = atoi( argv[2] ), h = atoi( argv[3] ), frames = atoi( argv[4] ), preset = atoi( argv[5] ), qp = atoi( argv[6] );
  const char* outPath = argv[7];
  if( argc > 8 )
  {
    if( b200Load( argv[8] ) || b200LoadAffine( argv[8] ) ) { fprintf( stderr, "cannot bind %s: %s\n", argv[8], g_b200.error.c_str() ); return 3; }
    if( argc > 9 && ( !strcmp( argv[9], "tu" ) || !strcmp( argv[9], "turdoq" ) || !strcmp( argv[9], "all" ) ) )
    {
      g_useRdoq = strcmp( argv[9], "tu" ) != 0;
      if( !strcmp( argv[9], "all" ) )
      {
        if( b200LoadMctf( argv[8] ) ) { fprintf( stderr, "cannot bind the MCTF entry points of %s: %s\n", argv[8], g_b200.error.c_str() ); return 3; }
        g_useMctf = true;
        g_realMctfErr = g_b200m.errorBatch; g_b200m.errorBatch = countingMctfErr;
      }
      if( b200LoadTu( argv[8] ) ) { fprintf( stderr, "cannot bind the TU entry points of %s: %s\n", argv[8], g_b200.error.c_str() ); return 3; }
      g_useTu = true;
    }
    g_useB200 = true;
    g_realDist = g_b200.distBlock; g_b200.distBlock = countingDist;
    g_realX5 = g_b200.sadX5; g_b200.sadX5 = countingX5;
  }
  vvenc_config cfg;
  vvenc_init_default( &cfg, w, h, 30, 0, qp, (vvencPresetMode) preset );
  cfg.m_numThreads = 1;                       // one worker: one vvb_ctx; the result of the reference does not depend on the thread count
  cfg.m_inputBitDepth[0] = 8; cfg.m_internalBitDepth[0] = 10;
  if( getenv( "VVB_ENC_TS" ) ) { cfg.m_TS = 1; cfg.m_useBDPCM = getenv( "VVB_ENC_BDPCM" ) ? 1 : 0; }      // transform skip always tried (the presets leave it to the screen-content detector): both arms of an identity run set it
  cfg.m_verbosity = VVENC_SILENT;
  vvenc_set_msg_callback( &cfg, nullptr, quietLog );
  vvencEncoder* enc = vvenc_encoder_create();
  if( !enc ) return 4;
  int rc = vvenc_encoder_open( enc, &cfg );
  if( rc ) { fprintf( stderr, "vvenc_encoder_open: %d %s\n", rc, vvenc_get_last_error( enc ) ); return 5; }

  FILE* fi = fopen( inPath, "rb" );
  if( !fi ) { fprintf( stderr, "cannot read %s\n", inPath ); return 6; }
  std::vector<uint8_t> out;
  vvencYUVBuffer* yuv = vvenc_YUVBuffer_alloc();
  vvenc_YUVBuffer_alloc_buffer( yuv, VVENC_CHROMA_420, w, h );
  vvencAccessUnit* au = vvenc_accessUnit_alloc();
  vvenc_accessUnit_alloc_payload( au, 2 * w * h + 65536 );
  std::vector<uint8_t> raw( (size_t) w * h * 3 / 2 );
  bool done = false;
  int fed = 0;
  try
  {
    while( !done )
    {
      vvencYUVBuffer* in = nullptr;
      if( fed < frames )
      {
        if( fread( raw.data(), 1, raw.size(), fi ) != raw.size() ) { fprintf( stderr, "short read at frame %d\n", fed ); return 7; }
        const uint8_t* p = raw.data();
        for( int c = 0; c < 3; c++ )
        {
          vvencYUVPlane& pl = yuv->planes[c];
          for( int y = 0; y < pl.height; y++ )
            for( int x = 0; x < pl.width; x++ ) pl.ptr[y * pl.stride + x] = *p++;
        }
        yuv->sequenceNumber = fed; yuv->cts = fed; yuv->ctsValid = true;
        in = yuv; fed++;
      }
      rc = vvenc_encode( enc, in, au, &done );
      if( rc ) { fprintf( stderr, "vvenc_encode: %d %s\n", rc, vvenc_get_last_error( enc ) ); return 8; }
      if( au->payloadUsedSize > 0 ) out.insert( out.end(), au->payload, au->payload + au->payloadUsedSize );
    }
  }
  catch( std::exception& e ) { fprintf( stderr, "exception: %s\n", e.what() ); return 9; }
  fclose( fi );
  vvenc_encoder_close( enc );
  FILE* fo = fopen( outPath, "wb" );
  if( fo ) { fwrite( out.data(), 1, out.size(), fo ); fclose( fo ); }
  uint64_t hsh = 1469598103934665603ull;
  for( uint8_t b : out ) { hsh ^= b; hsh *= 1099511628211ull; }
  printf( "ENC frames=%d bytes=%zu fnv1a=%016llx dist_calls=%llu x5_calls=%llu b200=%d rdcost_installs=%ld affine_installs=%ld tu_fwd=%llu tu_dq=%llu tu_rdoq=%llu tu_rdoq_ts=%llu tu_bdpcm=%llu tu_inv=%llu tu_inv_lfnst=%llu tu_ref=%llu mctf_installs=%ld mctf_calls=%llu\n", fed, out.size(),
          (unsigned long long) hsh, g_distCalls.load(), g_otherCalls.load(), g_useB200 ? 1 : 0, g_rdCostInstalls.load(), g_affineInstalls.load(), g_tuFwd.load(), g_tuDq.load(), g_tuRdoq.load(), g_tuRdoqTs.load(), g_tuBdpcm.load(), g_tuInv.load(),
          g_tuInvLfnst.load(), g_tuRef.load(), g_mctfInstalls.load(), g_mctfCalls.load() );
  return 0;
}
