// integration/MCTFB200.h -- reference-side binding of libvvenc_b200.so for the MCTF motion search (CommonLib/MCTF.cpp).
//
//   motionEstimationLumaB200  <->  MCTF::motionEstimationLuma (MCTF.cpp:1329-1397) -> estimateLumaLn (:1166-1327)
//   bilateralFilterB200       <->  MCTF::bilateralFilter (MCTF.cpp:1489-1556) -> xFinalizeBlkLine (:1399-1487): one vvb_mctf_apply per picture and component
//
// Same arguments as the member.  The per-block `error < best.error` chains of estimateLumaLn are kept; what changes is where the errors come from:
// motionErrorLuma (:1099-1164) is not called per candidate, the candidate sets of a whole picture go out as tables:
//
//   stage A  predictors of the coarser level (3x3 neighbourhood, scaled by `factor`) and the zero vector   vvb_mctf_error_batch, one call per neighbour  (:1191-1214)
//   stage B  integer grid around trunc(best / 16), range 8 / 5 / 3 / none                                  vvb_mctf_search_grid                         (:1216-1228)
//   stage C  doubleRes: +-12 (or +-6) step 4 (or 6), +-2 step 2, +-1 step 1 around the running best        vvb_mctf_search_grid, three calls            (:1229-1287)
//   stage D  final vectors of the upper and left neighbour: a true dependency, followed along              vvb_mctf_error_batch per anti-diagonal       (:1288-1306)
//            anti-diagonals (all blocks with the same bx + by at once)
//   stage E  doubleRes: variance for the error scaling                                                     vvb_mctf_calc_var                            (:1308-1321)
//
// Early exits of the reference's error kernels never change a decision (a partial sum is only returned once it exceeds the best), so full sums give the
// same field.  This is the C++ twin of vvenc_b200/mctf_host.py; tests/test_integration_host.py runs it next to the member itself.
// Include after RdCostB200.h / InterSearchB200.h and CommonLib/MCTF.h; private members of MCTF are read (m_searchPttrn, m_mctfUnitSize, ...): inside
// the encoder this is a member function.
#pragma once
#include <cmath>
#include <vector>
#include "InterSearchB200.h"

struct B200MctfApi
{
  bool bound = false;
  decltype( &vvb_mctf_error_batch )  errorBatch = nullptr;
  decltype( &vvb_mctf_search_grid )  searchGrid = nullptr;
  decltype( &vvb_mctf_calc_var )     calcVar = nullptr;
  decltype( &vvb_mctf_apply )        apply = nullptr;
} ;
static B200MctfApi g_b200m;

inline int b200LoadMctf( const char* libPath )
{
  if( g_b200m.bound ) return 0;
  int rc = b200LoadSearch( libPath );
  if( rc ) return rc;
  void* h = g_b200.handle;
#define VVB_RESOLVE( member, name ) g_b200m.member = (decltype( g_b200m.member )) dlsym( h, #name ); if( !g_b200m.member ) { g_b200.error = "missing " #name; return -2; }
  VVB_RESOLVE( errorBatch, vvb_mctf_error_batch )  VVB_RESOLVE( searchGrid, vvb_mctf_search_grid )  VVB_RESOLVE( calcVar, vvb_mctf_calc_var )
  VVB_RESOLVE( apply, vvb_mctf_apply )
#undef VVB_RESOLVE
  g_b200m.bound = true;
  return 0;
}

enum { B200_PLANE_MCTF_ORG = 12, B200_PLANE_MCTF_REF = 13 };

namespace b200mctf
{
struct Best { int x = 0, y = 0; int64_t error = INT_LEAST32_MAX; };               // MotionVector() starts at INT_LEAST32_MAX (MCTF.h:79)

struct Level
{
  int lowRes = 0;
  std::vector<vvb_mctf_cand> blocks;                                                // x, y, w, h of every block; vectors filled per call
  std::vector<Best>          best;

  std::vector<int32_t> errors( const std::vector<int>& sel, const std::vector<int>& mvx, const std::vector<int>& mvy ) const
  {
    std::vector<vvb_mctf_cand> c( sel.size() );
    for( size_t i = 0; i < sel.size(); i++ ) { c[i] = blocks[sel[i]]; c[i].mvx = mvx[i]; c[i].mvy = mvy[i]; }
    std::vector<int32_t> e( sel.size() );
    if( !sel.empty() ) b200Check( g_b200m.errorBatch( b200CtxOfThread(), B200_PLANE_MCTF_ORG, B200_PLANE_MCTF_REF, c.data(), (int) c.size(), lowRes, e.data() ) );
    return e;
  }

  // every block: candidates centre + (ox, oy) for ox, oy in the equally spaced list offs (1/16 pel), visited y outer / x inner, strictly smaller wins.
  // The grid entry point takes centre +- k * step with step <= 16, so wider or centre-less sets (e.g. {-6, 6} or 32-pel steps) are read out of the smallest
  // covering lattice.
  void gridRound( const std::vector<int>& cx, const std::vector<int>& cy, const std::vector<int>& offs, bool skipZero )
  {
    const int n = (int) blocks.size(), m = (int) offs.size();
    if( m == 1 )                                                                    // a single position: the candidate-list entry point
    {
      if( skipZero && offs[0] == 0 ) return;
      std::vector<int> sel( n ), mvx( n ), mvy( n );
      for( int i = 0; i < n; i++ ) { sel[i] = i; mvx[i] = cx[i] + offs[0]; mvy[i] = cy[i] + offs[0]; }
      const std::vector<int32_t> e = errors( sel, mvx, mvy );
      for( int i = 0; i < n; i++ ) if( e[i] < best[i].error ) { best[i].error = e[i]; best[i].x = mvx[i]; best[i].y = mvy[i]; }
      return;
    }
    int step = offs[1] - offs[0];
    while( step > 16 ) step /= 2;
    const int span = offs[m - 1] - offs[0];
    const int radius = ( span + 2 * step - 1 ) / ( 2 * step );
    const int shift = offs[0] + radius * step, side = 2 * radius + 1;
    std::vector<vvb_mctf_cand> c( blocks );
    for( int i = 0; i < n; i++ ) { c[i].mvx = cx[i] + shift; c[i].mvy = cy[i] + shift; }
    std::vector<int32_t> tab( (size_t) n * side * side );
    b200Check( g_b200m.searchGrid( b200CtxOfThread(), B200_PLANE_MCTF_ORG, B200_PLANE_MCTF_REF, c.data(), n, step, radius, lowRes, tab.data() ) );
    for( int i = 0; i < n; i++ )
      for( int j = 0; j < m; j++ )
        for( int k = 0; k < m; k++ )
        {
          if( skipZero && offs[j] == 0 && offs[k] == 0 ) continue;
          const int64_t e = tab[( (size_t) i * side + ( offs[j] - offs[0] ) / step ) * side + ( offs[k] - offs[0] ) / step];
          if( e < best[i].error ) { best[i].error = e; best[i].x = cx[i] + offs[k]; best[i].y = cy[i] + offs[j]; }
        }
  }
};
}   // namespace b200mctf

inline void motionEstimationLumaB200( const MCTF& m, Array2D<MotionVector>& mvs, const PelStorage& orig, const PelStorage& buffer, const int blockSize,
                                      const Array2D<MotionVector>* previous, const int factor, const bool doubleRes )
{
  using namespace b200mctf;
  const CPelBuf org = orig.Y(), buf = buffer.Y();
  const int width = org.width, height = org.height, bitDepth = m.m_encCfg->m_internalBitDepth[CH_L];
  const int pattern = m.m_searchPttrn, mvf = m.m_motionVectorFactor;
  vvb_ctx* ctx = b200CtxOfThread();
  b200Check( g_b200s.planeUpload( ctx, B200_PLANE_MCTF_ORG, org.buf, org.stride, width, height, 0, bitDepth ) );
  b200Check( g_b200s.planeUpload( ctx, B200_PLANE_MCTF_REF, buf.buf, buf.stride, width, height, MCTF_PADDING, bitDepth ) );       // picture buffers are padded (MCTF.cpp:1072-1097, Picture.cpp)

  Level lv; lv.lowRes = m.m_lowResFltSearch ? 1 : 0;
  int bxN = 0, byN = 0;
  for( int y = 0; y + 8 <= height; y += blockSize ) byN++;                         // `blockY + 8 <= origHeight` (:1388), `blockX + 8 <= origWidth` (:1174)
  for( int x = 0; x + 8 <= width; x += blockSize ) bxN++;
  const int n = bxN * byN;
  lv.blocks.resize( n ); lv.best.resize( n );
  for( int by = 0; by < byN; by++ )
    for( int bx = 0; bx < bxN; bx++ )
    {
      vvb_mctf_cand& c = lv.blocks[by * bxN + bx];
      c.x = bx * blockSize; c.y = by * blockSize; c.mvx = c.mvy = 0;
      c.w = (uint16_t)( std::min( blockSize, width - c.x ) & ~7 ); c.h = (uint16_t)( std::min( blockSize, height - c.y ) & ~7 );   // motionErrorLuma :1105-1106
    }
  std::vector<int> all( n ); for( int i = 0; i < n; i++ ) all[i] = i;

  // ---- stage A
  int range = doubleRes ? 0 : ( pattern == 2 ? 3 : 5 );
  if( !previous ) range = 8;
  else
  {
    for( int py = -1; py <= 1; py++ )
      for( int px = -1; px <= 1; px++ )
      {
        std::vector<int> sel, mvx, mvy;
        for( int i = 0; i < n; i++ )
        {
          const int testy = lv.blocks[i].y / ( 2 * blockSize ) + py, testx = lv.blocks[i].x / ( 2 * blockSize ) + px;
          if( testy < 0 || testy >= (int) previous->h() || testx < 0 || testx >= (int) previous->w() ) continue;
          const MotionVector& old = previous->get( testx, testy );
          sel.push_back( i ); mvx.push_back( old.x * factor ); mvy.push_back( old.y * factor );
        }
        const std::vector<int32_t> e = lv.errors( sel, mvx, mvy );
        for( size_t k = 0; k < sel.size(); k++ ) if( e[k] < lv.best[sel[k]].error ) { lv.best[sel[k]].error = e[k]; lv.best[sel[k]].x = mvx[k]; lv.best[sel[k]].y = mvy[k]; }
      }
    const std::vector<int> zero( n, 0 );
    const std::vector<int32_t> e = lv.errors( all, zero, zero );
    for( int i = 0; i < n; i++ ) if( e[i] < lv.best[i].error ) { lv.best[i].error = e[i]; lv.best[i].x = 0; lv.best[i].y = 0; }
  }

  // ---- stage B: integer grid around prevBest / m_motionVectorFactor (C division, truncation toward zero)
  std::vector<int> cx( n ), cy( n ), offs;
  {
    const int d = ( !previous && pattern == 2 ) ? 2 : 1;
    for( int i = 0; i < n; i++ ) { cx[i] = lv.best[i].x / mvf * mvf; cy[i] = lv.best[i].y / mvf * mvf; }
    for( int v = -range; v <= range; v += d ) offs.push_back( v * mvf );
    lv.gridRound( cx, cy, offs, false );
  }

  // ---- stage C
  if( doubleRes )
  {
    const int doubleRange = pattern ? 6 : 12, d1 = pattern == 2 ? 6 : 4;
    const int rounds[3][2] = { { doubleRange, d1 }, { 2, 2 }, { 1, 1 } };
    for( const auto& r : rounds )
    {
      offs.clear();
      for( int v = -r[0]; v <= r[0]; v += r[1] ) offs.push_back( v );
      for( int i = 0; i < n; i++ ) { cx[i] = lv.best[i].x; cy[i] = lv.best[i].y; }
      lv.gridRound( cx, cy, offs, true );
    }
  }

  // ---- stage D: anti-diagonal wavefront over the upper / left dependency
  for( int wave = 1; wave < bxN + byN - 1; wave++ )
  {
    for( int pass = 0; pass < 2; pass++ )                                          // above first, then left (:1288-1306)
    {
      std::vector<int> sel, mvx, mvy;
      for( int by = std::max( 0, wave - bxN + 1 ); by <= std::min( wave, byN - 1 ); by++ )
      {
        const int bx = wave - by;
        if( pass == 0 ? by == 0 : bx == 0 ) continue;
        const int src = pass == 0 ? ( by - 1 ) * bxN + bx : by * bxN + bx - 1;
        sel.push_back( by * bxN + bx ); mvx.push_back( lv.best[src].x ); mvy.push_back( lv.best[src].y );
      }
      const std::vector<int32_t> e = lv.errors( sel, mvx, mvy );
      for( size_t k = 0; k < sel.size(); k++ ) if( e[k] < lv.best[sel[k]].error ) { lv.best[sel[k]].error = e[k]; lv.best[sel[k]].x = mvx[k]; lv.best[sel[k]].y = mvy[k]; }
    }
  }

  // ---- stage E and write-back
  std::vector<double> var;
  if( doubleRes )
  {
    var.resize( n );
    b200Check( g_b200m.calcVar( ctx, B200_PLANE_MCTF_ORG, lv.blocks.data(), n, var.data() ) );
  }
  for( int i = 0; i < n; i++ )
  {
    MotionVector best;
    best.set( lv.best[i].x, lv.best[i].y, (int) lv.best[i].error );
    if( doubleRes )
    {
      const int w = lv.blocks[i].w, h = lv.blocks[i].h;
      const double bdScale = double( 1 << ( 2 * ( 10 - bitDepth ) ) );
      const double dvar = var[i] * bdScale;
      const double mse  = best.error * bdScale / double( w * h );
      best.error   = (int)( 20 * ( ( best.error * bdScale + 5.0 ) / ( dvar + 5.0 ) ) + mse / 50.0 );
      best.rmsme   = uint16_t( 0.5 + sqrt( mse ) );
      best.overlap = ( (double) w * h ) / ( m.m_mctfUnitSize * m.m_mctfUnitSize );
    }
    mvs.get( i % bxN, i / bxN ) = best;
  }
}

// MCTF::bilateralFilter: the parameters xFinalizeBlkLine derives per block (filter set, planar-correction switch, reference strengths, weight scaling, sigma^2) are
// derived once per component, the motion fields of all neighbour pictures go out with one call per component and the filtered plane comes back whole.
// Chroma (4:2:0 / 4:2:2 / 4:4:4) runs through the same entry point: units shrink by the sub-sampling, and because xFinalizeBlkLine uses the vector as
// dx = mv.x >> csx (phase dx & 15, integer part mv.x >> (4 + csx), :1450-1453) the vectors are handed over pre-shifted.
enum { B200_PLANE_MCTF_SRC0 = 20 };
inline void bilateralFilterB200( const MCTF& m, const PelStorage& orgPic, std::deque<TemporalFilterSourcePicInfo>& srcFrameInfo, PelStorage& newOrgPic, double overallStrength )
{
  const int numRefs = (int) srcFrameInfo.size();
  if( numRefs < 1 || numRefs > 8 ) THROW( "vvb_mctf_apply takes 1..8 neighbour pictures" );
  const VVEncCfg& cfg = *m.m_encCfg;
  const ChromaFormat chFmt = cfg.m_internChromaFormat == VVENC_CHROMA_400 ? CHROMA_400 : cfg.m_internChromaFormat == VVENC_CHROMA_420 ? CHROMA_420
                           : cfg.m_internChromaFormat == VVENC_CHROMA_422 ? CHROMA_422 : CHROMA_444;
  const int unit = m.m_mctfUnitSize;
  const int refStrengthRow = cfg.m_picReordering ? 0 : 1;                                      // :1405
  const int bxN = ( orgPic.Y().width + unit - 1 ) / unit, byN = ( orgPic.Y().height + unit - 1 ) / unit;
  const double lumaSigmaSq = m.m_sigmaMultiplier * ( 128.0 + 3.0 / 256.0 * cfg.m_QP * cfg.m_QP * cfg.m_QP ), chromaSigmaSq = 30 * 30;       // :1491-1492
  vvb_ctx* ctx = b200CtxOfThread();
  std::vector<vvb_mctf_mv> mvs( (size_t) numRefs * bxN * byN );

  for( int c = 0; c < (int) getNumberValidComponents( chFmt ); c++ )
  {
    const ComponentID compID = ComponentID( c );
    const ChannelType ch = toChannelType( compID );
    const int csx = getComponentScaleX( compID, chFmt ), csy = getComponentScaleY( compID, chFmt );
    if( csx != csy ) THROW( "vvb_mctf_apply takes square units: 4:2:2 chroma stays on the host" );
    const int bitDepth = cfg.m_internalBitDepth[ch];
    const CPelBuf org = orgPic.bufs[c];
    b200Check( g_b200s.planeUpload( ctx, B200_PLANE_MCTF_ORG, org.buf, org.stride, org.width, org.height, 0, bitDepth ) );

    vvb_mctf_apply_par par = {};
    par.num_refs = numRefs; par.block_size = unit >> csx;                                      // blkSizeX / blkSizeY (:1420-1421)
    par.low_res_filter = m.m_lowResFltApply ? 1 : 0;                                           // m_interpolationFilter4 instead of ..8 (:1459)
    par.planar_correction = cfg.m_QP <= 32 ? 1 : 0;                                            // :1474 (the rmsme / shape conditions are per block, inside)
    par.weight_scaling = overallStrength * ( isChroma( compID ) ? m.m_chromaFactor : 0.4 );    // :1417
    const double bitDepthDiffWeighting = 1024.0 / ( ( ( 1 << bitDepth ) - 1 ) + 1 );
    par.sigma_sq = ( isChroma( ch ) ? chromaSigmaSq : lumaSigmaSq ) / ( bitDepthDiffWeighting * bitDepthDiffWeighting );     // :1500
    for( int i = 0; i < numRefs; i++ )
    {
      const CPelBuf src = srcFrameInfo[i].picBuffer.bufs[c];
      b200Check( g_b200s.planeUpload( ctx, B200_PLANE_MCTF_SRC0 + i, src.buf, src.stride, org.width, org.height, MCTF_PADDING >> csx, bitDepth ) );
      par.ref_plane[i] = B200_PLANE_MCTF_SRC0 + i;
      par.ref_strength[i] = m.m_refStrengths[refStrengthRow][srcFrameInfo[i].index];          // :1480
      for( int by = 0; by < byN; by++ )
        for( int bx = 0; bx < bxN; bx++ )
        {
          const MotionVector& v = srcFrameInfo[i].mvs.get( bx, by );
          vvb_mctf_mv& o = mvs[( (size_t) i * byN + by ) * bxN + bx];
          o.x = v.x >> csx; o.y = v.y >> csy; o.error = v.error; o.rmsme = v.rmsme; o.pad = 0;
        }
    }
    PelBuf dst = newOrgPic.bufs[c];
    b200Check( g_b200m.apply( ctx, B200_PLANE_MCTF_ORG, &par, mvs.data(), dst.buf, dst.stride ) );
  }
}

// ---- per-call form (verification shape, like the FpDistFunc trampolines of RdCostB200.h): the MCTF error pointers answer from the library one call at a time, so that
// the UNMODIFIED MCTF::motionEstimationLuma / estimateLumaLn control runs on top of them inside the live encoder (oracle/enc_identity.cpp ... all).
//   m_motionErrorLumaInt8 / IntX          (MCTF.h:160-161; motionErrorLumaInt, MCTF.cpp:122-145)
//   m_motionErrorLumaFrac8 / FracX [0|1]  (MCTF.h:163-164; motionErrorLumaFrac6 / Frac4, :147-257) -- the filter rows identify the 1/16-pel phase
//   m_calcVar                             (MCTF.h:170; calcVarCore, :520-546)
// Each call uploads the original block and the reference window it reads (two rows / columns before, three after: the 6-tap support) as two small planes and asks for
// one candidate.  The full sum comes back where the member may have stopped early; a stopped sum is only ever compared with the running best it already exceeds.
enum { B200_PLANE_MCTF_CALL_ORG = 16, B200_PLANE_MCTF_CALL_REF = 17 };
inline int b200MctfErrorCall( const Pel* org, const ptrdiff_t origStride, const Pel* buf, const ptrdiff_t buffStride, const int w, const int h, const int fx, const int fy, const int tap4, const int bitDepth )
{
  vvb_ctx* ctx = b200CtxOfThread();
  b200Check( g_b200s.planeUpload( ctx, B200_PLANE_MCTF_CALL_ORG, org, (int) origStride, w, h, 0, bitDepth ) );
  b200Check( g_b200s.planeUpload( ctx, B200_PLANE_MCTF_CALL_REF, buf, (int) buffStride, w, h, 4, bitDepth ) );
  vvb_mctf_cand c; c.x = 0; c.y = 0; c.mvx = fx; c.mvy = fy; c.w = (uint16_t) w; c.h = (uint16_t) h;
  int32_t e = 0;
  b200Check( g_b200m.errorBatch( ctx, B200_PLANE_MCTF_CALL_ORG, B200_PLANE_MCTF_CALL_REF, &c, 1, tap4, &e ) );
  return e;
}
inline int b200MctfErrInt( const Pel* org, const ptrdiff_t so, const Pel* buf, const ptrdiff_t sb, const int w, const int h, const int )
{
  return b200MctfErrorCall( org, so, buf, sb, w, h, 0, 0, 0, 10 );               // integer position: no filter, no clipping -- the bit depth does not enter
}
inline int b200MctfErrFrac6( const Pel* org, const ptrdiff_t so, const Pel* buf, const ptrdiff_t sb, const int w, const int h, const int16_t* xFilter, const int16_t* yFilter, const int bitDepth, const int )
{
  const int fx = int( ( xFilter - &MCTF::m_interpolationFilter8[0][0] ) / 8 ), fy = int( ( yFilter - &MCTF::m_interpolationFilter8[0][0] ) / 8 );
  if( fx < 0 || fx > 15 || fy < 0 || fy > 15 ) THROW( "filter row outside m_interpolationFilter8" );
  return b200MctfErrorCall( org, so, buf, sb, w, h, fx, fy, 0, bitDepth );
}
inline int b200MctfErrFrac4( const Pel* org, const ptrdiff_t so, const Pel* buf, const ptrdiff_t sb, const int w, const int h, const int16_t* xFilter, const int16_t* yFilter, const int bitDepth, const int )
{
  const int fx = int( ( xFilter - &MCTF::m_interpolationFilter4[0][0] ) / 4 ), fy = int( ( yFilter - &MCTF::m_interpolationFilter4[0][0] ) / 4 );
  if( fx < 0 || fx > 15 || fy < 0 || fy > 15 ) THROW( "filter row outside m_interpolationFilter4" );
  return b200MctfErrorCall( org, so, buf, sb, w, h, fx, fy, 1, bitDepth );
}
inline double b200MctfCalcVar( const Pel* org, const ptrdiff_t so, const int w, const int h )
{
  vvb_ctx* ctx = b200CtxOfThread();
  b200Check( g_b200s.planeUpload( ctx, B200_PLANE_MCTF_CALL_ORG, org, (int) so, w, h, 0, 10 ) );
  vvb_mctf_cand c; c.x = 0; c.y = 0; c.mvx = 0; c.mvy = 0; c.w = (uint16_t) w; c.h = (uint16_t) h;
  double v = 0.0;
  b200Check( g_b200m.calcVar( ctx, B200_PLANE_MCTF_CALL_ORG, &c, 1, &v ) );
  return v;
}
// the MCTF::_initMCTFB200() a maintainer would add next to _initMCTF_X86 (x86/MCTFX86.h:1491-1506); the apply-stage pointers keep their x86 kernels
// (the library offers that stage per picture: bilateralFilterB200 above)
inline void installB200( MCTF& m )
{
  m.m_motionErrorLumaInt8 = b200MctfErrInt;        m.m_motionErrorLumaIntX = b200MctfErrInt;
  m.m_motionErrorLumaFrac8[0] = b200MctfErrFrac6;  m.m_motionErrorLumaFracX[0] = b200MctfErrFrac6;
  m.m_motionErrorLumaFrac8[1] = b200MctfErrFrac4;  m.m_motionErrorLumaFracX[1] = b200MctfErrFrac4;
  m.m_calcVar = b200MctfCalcVar;
}
