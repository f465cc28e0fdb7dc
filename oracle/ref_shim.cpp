/*
 * oracle/ref_shim.cpp -- TEST INFRASTRUCTURE, not product code.
 *
 * Thin extern "C" probe around the UNMODIFIED reference (fraunhoferhhi/vvenc) library objects that
 * oracle/Makefile.ref compiles in place from /root/reference.  It lets tests/ (ctypes) and bench.py's
 * cpu_baseline / --impl reference legs call the reference's own scalar and SSE4.1/AVX2 kernels through the
 * very function-pointer tables the encoder uses:
 *
 *   RdCost::m_afpDistortFunc / m_afpDistortFuncX5 / m_fxdWtdPredPtr  (CommonLib/RdCost.h:117-121)
 *   TrQuant::xT + Quant::quant / Quant::xNeedRDOQ                    (CommonLib/TrQuant.cpp:481, Quant.cpp:735,835)
 *   MCTF::m_motionErrorLumaInt8 / m_motionErrorLumaFrac8[2]          (CommonLib/MCTF.h:160-166)
 *   AffineGradientSearch::m_*                                        (CommonLib/AffineGradientSearch.h:67-69)
 *
 * Nothing here re-implements codec arithmetic: every number returned is computed by reference code.
 * The only logic of our own is argument marshalling, the ME full-search loop of
 * InterSearch::xPatternSearch (EncoderLib/InterSearch.cpp:2209-2251; that member is private and bound to
 * encoder state, so its 20-line loop is replayed here on top of the reference distFunc + MV-cost calls),
 * and a std::thread fan-out used for the multi-core CPU baseline.
 *
 * "private/protected" are opened up below ONLY so that the probe can reach TrQuant::xT, Quant::xQuant,
 * Quant::xNeedRDOQ and TransformUnit::m_coeffs; object layout is unaffected.
 */
#include <cstdint>
#include <cstring>
#include <cstdlib>
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
#include <cstdio>
#include <immintrin.h>

#define private public
#define protected public
#include "vvenc/vvenc.h"
#include "CommonLib/CommonDef.h"
#include "CommonLib/Unit.h"
#include "CommonLib/UnitTools.h"
#include "CommonLib/Slice.h"
#include "CommonLib/CodingStructure.h"
#include "CommonLib/RdCost.h"
#include "CommonLib/TrQuant.h"
#include "CommonLib/TrQuant_EMT.h"
#include "CommonLib/Quant.h"
#include "CommonLib/DepQuant.h"
#include "CommonLib/MCTF.h"
#include "CommonLib/AffineGradientSearch.h"
#include "CommonLib/InterpolationFilter.h"
#include "EncoderLib/InterSearch.h"
#include "CommonLib/Rom.h"
#include "CommonLib/Contexts.h"
#undef private
#undef protected

using namespace vvenc;

namespace {

void createRd( RdCost& rc, int opt );

struct RefCtx
{
  RdCost                 rdScalar, rdSimd, rdB200;
  bool                   b200Ready = false;
  RdCost& rd( int opt ) { if( opt == 2 && !b200Ready ) { createRd( rdB200, 2 ); b200Ready = true; } return opt == 2 ? rdB200 : opt ? rdSimd : rdScalar; }
  TrQuant*               tq      = nullptr;
  MCTF*                  mctf[2] = { nullptr, nullptr };
  AffineGradientSearch*  ags[3]  = { nullptr, nullptr, nullptr };   // [2]: pointers patched by installB200( AffineGradientSearch& ), created on demand
  std::string            simd;
};

RefCtx* g_ctx = nullptr;
std::mutex g_mtx;

RefCtx& ctx()
{
  if( !g_ctx )
  {
    std::lock_guard<std::mutex> lk( g_mtx );
    if( !g_ctx )
    {
      RefCtx* c = new RefCtx;
      c->rdScalar.create( false );
      c->rdSimd  .create( true );
      c->tq      = new TrQuant;
      c->tq->init( nullptr, 0, false, false, true, 8 );   // rdoq off, thrVal 8 (vvencCfg.cpp:971-973)
      c->mctf[0] = new MCTF( false );
      c->mctf[1] = new MCTF( true );
      c->ags[0]  = new AffineGradientSearch( false );
      c->ags[1]  = new AffineGradientSearch( true );
      g_ctx = c;
    }
  }
  return *g_ctx;
}

// TrQuant keeps scratch buffers (m_plTempCoeff, m_blk, m_tmp), so every calling thread needs its own instance -- and building one costs about a millisecond
// (quantiser tables), which would dominate the batched probes on short-lived worker threads.  Instances are therefore leased from a pool: a thread takes one on
// first use and hands it back when it exits; the pool is dropped when refshim_set_simd() replaced the context (per-instance quantiser pointers are re-resolved).
struct TqPool
{
  std::mutex            mtx;
  std::vector<TrQuant*> idle;
  RefCtx*               owner = nullptr;
};
TqPool g_tqPool;

struct TqLease
{
  TrQuant* t = nullptr;
  RefCtx*  owner = nullptr;
  TrQuant& get()
  {
    RefCtx& c = ctx();
    if( owner != &c )
    {
      release();
      {
        std::lock_guard<std::mutex> lk( g_tqPool.mtx );
        if( g_tqPool.owner != &c ) { g_tqPool.idle.clear(); g_tqPool.owner = &c; }       // instances of the previous SIMD mode are leaked on purpose (tiny, rare)
        if( !g_tqPool.idle.empty() ) { t = g_tqPool.idle.back(); g_tqPool.idle.pop_back(); }
      }
      if( !t ) { t = new TrQuant; t->init( nullptr, 0, false, false, true, 8 ); }        // rdoq off, thrVal 8 (vvencCfg.cpp:971-973)
      owner = &c;
    }
    return *t;
  }
  void release()
  {
    if( !t ) return;
    std::lock_guard<std::mutex> lk( g_tqPool.mtx );
    if( g_tqPool.owner == owner ) g_tqPool.idle.push_back( t );
    t = nullptr; owner = nullptr;
  }
  ~TqLease() { release(); }
};
TrQuant& tqOfThread()
{
  static thread_local TqLease lease;
  return lease.get();
}

inline int ilog2( unsigned v ) { int r = 0; while( v > 1 ) { v >>= 1; r++; } return r; }

// DFunc family (ours) -> reference table base (CommonLib/TypeDef.h:339-382)
inline int familyBase( int family )
{
  switch( family )
  {
    case 0: return DF_SSE;
    case 1: return DF_SAD;
    case 2: return DF_HAD;
    case 3: return DF_HAD_fast;
    case 4: return DF_HAD_2SAD;
    default: return -1;
  }
}

inline FpDistFunc pickFunc( RdCost& rc, int family, int w, int bitDepth )
{
  const int row  = bitDepth > 10 ? 1 : 0;                      // RdCost.cpp:176,219
  const int base = familyBase( family );
  const int idx  = family == 4 ? base : base + ilog2( w );     // RdCost.cpp:211-215 (DF_HAD_2SAD has one slot)
  return rc.m_afpDistortFunc[row][idx];
}

inline uint64_t callDist( RdCost& rc, int family, const int16_t* org, int so, const int16_t* cur, int sc, int w, int h, int bitDepth, int subShift )
{
  DistParam dp;
  dp.org.buf = org; dp.org.stride = so; dp.org.width = w; dp.org.height = h;
  dp.cur.buf = cur; dp.cur.stride = sc; dp.cur.width = w; dp.cur.height = h;
  dp.bitDepth = bitDepth;
  dp.subShift = subShift;
  dp.compID   = COMP_Y;
  dp.applyWeight = false;
  dp.maximumDistortionForEarlyExit = MAX_DISTORTION;
  dp.distFunc = pickFunc( rc, family, w, bitDepth );
  return dp.distFunc( dp );
}

// ---------------------------------------------------------------------------------------------------------------
// Drop-in demo (compiled for real): the reference's own RdCost function-pointer tables
// (RdCost.h:117-121) are overwritten with trampolines into libvvenc_b200.so -- exactly what a maintainer-written
// RdCost::_initRdCostB200() would do.  opt == 2 in the probes below selects an RdCost patched this way, so the tests can
// drive UNMODIFIED reference call sites (DistParam + distFunc, dmvrSadX5, m_fxdWtdPredPtr, the xPatternSearch replay)
// with the GPU library underneath and compare against the AVX2 table.
#include "../integration/RdCostB200.h"
#include "../integration/InterSearchB200.h"
#include "../integration/MCTFB200.h"
#include "../integration/TrQuantB200.h"
#include "../integration/AffineGradientB200.h"   // the reference-side binding (B200Api, trampolines, installB200), compiled for real here

void createRd( RdCost& rc, int opt )     // 0 scalar, 1 SIMD, 2 SIMD table patched with the B200 trampolines
{
  rc.create( opt != 0 );
  if( opt == 2 ) { if( !g_b200.handle ) THROW( "refshim_install_b200 not called" ); installB200( rc ); }
}

template<class F>
void parallelFor( int n, int nthreads, F&& f )
{
  if( nthreads <= 1 || n < 2 * nthreads ) { f( 0, n, 0 ); return; }
  std::vector<std::thread> th;
  const int chunk = ( n + nthreads - 1 ) / nthreads;
  for( int t = 0; t < nthreads; t++ )
  {
    const int b = t * chunk, e = std::min( n, b + chunk );
    if( b >= e ) break;
    th.emplace_back( [=,&f]{ f( b, e, t ); } );
  }
  for( auto& t : th ) t.join();
}

static thread_local bool t_rigSignHiding = false;     // slice->signDataHidingEnabled of the next TuRig::setup (refshim_*_sdh entries)
// A minimal TU that satisfies everything TrQuant::xT / Quant::quant / Quant::xNeedRDOQ dereference.
struct TuRig
{
  XUCache          xuCache;
  std::mutex       csMutex;
  CodingStructure  cs;
  SPS              sps;
  PPS              pps;
  Slice            slice;
  CodingUnit       cu;
  TransformUnit    tu;
  std::vector<TCoeffSig> qcoef;
  TCoeffSig*       coeffPtrs[MAX_NUM_TBLOCKS];

  TuRig() : cs( xuCache, &csMutex ) {}

  void setup( int w, int h, int bitDepth, int mtsIdx, bool intraSlice, bool intraCu, int qp, ChromaFormat fmt = CHROMA_400 )
  {
    sps.bitDepths.recon[CH_L] = bitDepth;
    sps.bitDepths.recon[CH_C] = bitDepth;
    sps.qpBDOffset[CH_L] = 6 * ( bitDepth - 8 );
    sps.qpBDOffset[CH_C] = 6 * ( bitDepth - 8 );
    sps.internalMinusInputBitDepth[CH_L] = 0;
    sps.internalMinusInputBitDepth[CH_C] = 0;
    sps.MTS = true; sps.MTSIntra = true; sps.MTSInter = true; sps.LFNST = false;
    sps.chromaFormatIdc = fmt;
    slice.sps = &sps; slice.pps = &pps;
    slice.sliceType = intraSlice ? VVENC_I_SLICE : VVENC_B_SLICE;
    slice.nalUnitType = intraSlice ? VVENC_NAL_UNIT_CODED_SLICE_IDR_W_RADL : VVENC_NAL_UNIT_CODED_SLICE_TRAIL;
    slice.signDataHidingEnabled = t_rigSignHiding;
    slice.depQuantEnabled = false;
    slice.tsResidualCodingDisabled = false;
    cs.sps = &sps; cs.pps = &pps; cs.slice = &slice;

    const UnitArea ua( fmt, Area( 0, 0, w, h ) );
    static_cast<UnitArea&>( cu ) = ua;
    cu.cs = &cs; cu.slice = &slice; cu.chType = CH_L;
    cu.predMode = intraCu ? MODE_INTRA : MODE_INTER;
    cu.qp = qp; cu.lfnstIdx = 0; cu.ispMode = 0; cu.mipFlag = false; cu.sbtInfo = 0;
    cu.bdpcmM[CH_L] = 0; cu.bdpcmM[CH_C] = 0; cu.colorTransform = false; cu.chromaQpAdj = 0;
    cu.treeType = TREE_D; cu.modeType = MODE_TYPE_ALL;

    static_cast<UnitArea&>( tu ) = ua;
    tu.cu = &cu; tu.cs = &cs; tu.chType = CH_L; tu.depth = 0; tu.noResidual = false; tu.jointCbCr = 0;
    for( int i = 0; i < MAX_NUM_TBLOCKS; i++ ) { tu.mtsIdx[i] = 0; tu.cbf[i] = 0; tu.lastPos[i] = -1; }
    tu.mtsIdx[COMP_Y] = (uint8_t) mtsIdx;
    tu.next = tu.prev = nullptr; tu.idx = 0; tu.chromaAdj = 0;
    qcoef.assign( (size_t) w * h, 0 );
    for( int i = 0; i < MAX_NUM_TBLOCKS; i++ ) coeffPtrs[i] = qcoef.data();
    tu.init( coeffPtrs );
  }
};

// the rig (CodingStructure, XUCache, ...) is expensive to build: leased from a pool like the TrQuant instances, handed back when the calling thread exits
struct RigPool { std::mutex mtx; std::vector<TuRig*> idle; };
RigPool g_rigPool;
struct RigLease
{
  TuRig* r = nullptr;
  TuRig& get()
  {
    if( !r )
    {
      { std::lock_guard<std::mutex> lk( g_rigPool.mtx ); if( !g_rigPool.idle.empty() ) { r = g_rigPool.idle.back(); g_rigPool.idle.pop_back(); } }
      if( !r ) r = new TuRig;
    }
    return *r;
  }
  ~RigLease() { if( r ) { std::lock_guard<std::mutex> lk( g_rigPool.mtx ); g_rigPool.idle.push_back( r ); } }
};
TuRig& rig() { static thread_local RigLease lease; return lease.get(); }

int mtsIdxFor( int trHor, int trVer )   // ours: 0 DCT2, 1 DCT8, 2 DST7 (reference enum TransType, TypeDef.h)
{
  if( trHor == 0 && trVer == 0 ) return MTS_DCT2_DCT2;
  if( trHor == 0 || trVer == 0 ) return -1;          // mixed DCT2/MTS only arises via implicit MTS / SBT size rules
  // TrQuant.cpp:470-475: indHor = (mtsIdx-2)&1 -> DCT8, indVer = (mtsIdx-2)>>1 -> DCT8
  return MTS_DST7_DST7 + ( trHor == 1 ? 1 : 0 ) + ( trVer == 1 ? 2 : 0 );
}

} // namespace

extern "C" {

int refshim_version() { return 4; }

// Drop-in demo: load libvvenc_b200.so (path given by the test) and remember its entry points; opt == 2 then patches them into RdCost.
int refshim_install_b200( const char* libPath )
{
  std::lock_guard<std::mutex> lk( g_mtx );
  return b200Load( libPath );
}
const char* refshim_b200_error() { return g_b200.error.c_str(); }
uint64_t refshim_b200_launches()   // kernels launched by the calling thread's drop-in context (proof that the numbers came from the GPU)
{
  uint64_t n = 0;
  if( t_b200ctx && g_b200.launchCount ) g_b200.launchCount( t_b200ctx, &n );
  return n;
}

// "SCALAR" | "SSE41" | "SSE42" | "AVX" | "AVX2"; rebuilds every probe object so that the per-instance pointers
// are re-resolved (vvenc.cpp:412 -> VVEncImpl::setSIMDExtension, vvencimpl.cpp:800-866).
const char* refshim_set_simd( const char* name )
{
  std::lock_guard<std::mutex> lk( g_mtx );
  const char* r = vvenc_set_SIMD_extension( name );
  if( g_ctx ) { /* leak the tiny old ctx on purpose: callers may still hold pointers */ g_ctx = nullptr; }
  return r;
}

// opt: 0 = RdCost::create(false) scalar table, 1 = create(true) (SIMD per vvenc_set_SIMD_extension)
uint64_t refshim_dist( int opt, int family, const int16_t* org, int orgStride, const int16_t* cur, int curStride,
                       int w, int h, int bitDepth, int subShift )
{
  RefCtx& c = ctx();
  return callDist( c.rd( opt ), family, org, orgStride, cur, curStride, w, h, bitDepth, subShift );
}

// Pair list over two planes: desc[i] = { org_x, org_y, cur_x, cur_y, w, h } (plane coordinates, may be negative inside the margin).
void refshim_dist_list( int opt, int family, const int16_t* orgPlane, int orgStride, const int16_t* curPlane, int curStride,
                        const int32_t* desc, int n, int bitDepth, int subShift, uint64_t* out, int nthreads )
{
  RefCtx& c = ctx();
  parallelFor( n, nthreads, [&]( int b, int e, int )
  {
    RdCost rc; createRd( rc, opt );     // one RdCost per worker, as EncSlice.cpp:142-147 does
    for( int i = b; i < e; i++ )
    {
      const int32_t* d = desc + 6 * (size_t) i;
      const int16_t* o = orgPlane + (ptrdiff_t) d[1] * orgStride + d[0];
      const int16_t* u = curPlane + (ptrdiff_t) d[3] * curStride + d[2];
      out[i] = callDist( rc, family, o, orgStride, u, curStride, d[4], d[5], bitDepth, subShift );
    }
  } );
  (void) c;
}

uint64_t refshim_sad_mask( int opt, const int16_t* org, int orgStride, const int16_t* cur, int curStride, int w, int h,
                           const int16_t* mask, int maskStride, int stepX, int maskStride2, int bitDepth, int subShift )
{
  RefCtx& c = ctx();
  RdCost& rc = c.rd( opt );
  DistParam dp;
  dp.org.buf = org; dp.org.stride = orgStride; dp.org.width = w; dp.org.height = h;
  dp.cur.buf = cur; dp.cur.stride = curStride; dp.cur.width = w; dp.cur.height = h;
  dp.mask = mask; dp.maskStride = maskStride; dp.stepX = stepX; dp.maskStride2 = maskStride2;
  dp.bitDepth = bitDepth; dp.subShift = subShift; dp.compID = COMP_Y;
  return rc.m_afpDistortFunc[0][DF_SAD_WITH_MASK]( dp );
}

void refshim_sad_x5( int opt, const int16_t* org, int orgStride, const int16_t* cur, int curStride, int w, int h,
                     int bitDepth, int subShift, int calcCentre, uint64_t* cost5 )
{
  RefCtx& c = ctx();
  RdCost& rc = c.rd( opt );
  DistParam dp = rc.setDistParam( org, cur, orgStride, curStride, bitDepth, COMP_Y, w, h, subShift, true );   // RdCost.cpp:228
  Distortion tmp[5] = { 0, 0, 0, 0, 0 };
  dp.dmvrSadX5( dp, tmp, calcCentre != 0 );
  for( int i = 0; i < 5; i++ ) cost5[i] = tmp[i];
}

uint64_t refshim_fix_wsse( int opt, const int16_t* org, int orgStride, const int16_t* cur, int curStride, int w, int h,
                           int bitDepth, uint32_t fixedWeight )
{
  RefCtx& c = ctx();
  RdCost& rc = c.rd( opt );
  DistParam dp;
  dp.org.buf = org; dp.org.stride = orgStride; dp.org.width = w; dp.org.height = h;
  dp.cur.buf = cur; dp.cur.stride = curStride; dp.cur.width = w; dp.cur.height = h;
  dp.bitDepth = bitDepth; dp.compID = COMP_Y;
  return rc.m_fxdWtdPredPtr( dp, fixedWeight );
}

// MV rate (RdCost.h:181-203).  lambda is what RdCost::setLambda receives; motion lambda = sqrt(lambda).
uint32_t refshim_mv_bits( int x, int y, int predHor, int predVer, int costScale, int imvShift )
{
  RdCost rc; rc.create( false );
  rc.setPredictor( Mv( predHor, predVer ) );
  rc.setCostScale( costScale );
  return rc.getBitsOfVectorWithPredictor( x, y, imvShift );
}

uint64_t refshim_mv_cost( double lambda, int x, int y, int predHor, int predVer, int costScale, int imvShift )
{
  RdCost rc; rc.create( false );
  BitDepths bd; bd.recon[CH_L] = 10; bd.recon[CH_C] = 10;
  rc.setLambda( lambda, bd );
  rc.selectMotionLambda();
  rc.setPredictor( Mv( predHor, predVer ) );
  rc.setCostScale( costScale );
  return rc.getCostOfVectorWithPredictor( x, y, imvShift );
}

// Replay of InterSearch::xPatternSearch (InterSearch.cpp:2209-2251) for a list of blocks:
// blk[i] = { x, y, w, h, left, right, top, bottom, predHor, predVer } ; search positions are integer offsets (dx,dy)
// in [left..right]x[top..bottom] relative to the block position, cost = SAD(subShift) + floor(sqrt(lambda)*bits).
// out[i] = { bestDx, bestDy, bestCost(lo32), bestCost(hi32) } ; optional full cost table (uint32 SAD only) per block.
void refshim_full_search( int opt, const int16_t* orgPlane, int orgStride, const int16_t* refPlane, int refStride,
                          const int32_t* blk, int n, int bitDepth, int subShift, double lambda, int costScale, int imvShift,
                          int32_t* out, uint32_t* sadTables, int tableStride, int nthreads, int earlyExit )
{
  // earlyExit != 0 reproduces m_cDistParam.maximumDistortionForEarlyExit = uiSad (InterSearch.cpp:2241): the reference's SIMD SAD may
  // then return a partial sum for W >= 64 (RdCostX86.h:372-410) -- same decisions, used by the timed CPU baseline only.
  parallelFor( n, nthreads, [&]( int b, int e, int )
  {
    RdCost rc; createRd( rc, opt );
    BitDepths bd; bd.recon[CH_L] = bitDepth; bd.recon[CH_C] = bitDepth;
    rc.setLambda( lambda, bd );
    rc.selectMotionLambda();
    rc.setCostScale( costScale );
    for( int i = b; i < e; i++ )
    {
      const int32_t* d = blk + 10 * (size_t) i;
      const int x = d[0], y = d[1], w = d[2], h = d[3], l = d[4], r = d[5], t = d[6], bt = d[7];
      rc.setPredictor( Mv( d[8], d[9] ) );
      DistParam dp;
      dp.org.buf = orgPlane + (ptrdiff_t) y * orgStride + x; dp.org.stride = orgStride; dp.org.width = w; dp.org.height = h;
      dp.cur.stride = refStride; dp.cur.width = w; dp.cur.height = h;
      dp.bitDepth = bitDepth; dp.subShift = subShift; dp.compID = COMP_Y;
      dp.maximumDistortionForEarlyExit = MAX_DISTORTION;       // GPU always returns the full sum
      dp.distFunc = pickFunc( rc, 1, w, bitDepth );
      Distortion best = MAX_DISTORTION; int bx = 0, by = 0;
      uint32_t* tab = sadTables ? sadTables + (size_t) i * tableStride : nullptr;
      int k = 0;
      for( int dy = t; dy <= bt; dy++ )
      {
        const int16_t* row = refPlane + (ptrdiff_t)( y + dy ) * refStride + x;
        for( int dx = l; dx <= r; dx++, k++ )
        {
          dp.cur.buf = row + dx;
          Distortion sad = dp.distFunc( dp );
          if( tab ) tab[k] = (uint32_t) sad;
          sad += rc.getCostOfVectorWithPredictor( dx, dy, imvShift );
          if( sad < best ) { best = sad; bx = dx; by = dy; if( earlyExit ) dp.maximumDistortionForEarlyExit = sad; }
        }
      }
      out[4*i+0] = bx; out[4*i+1] = by; out[4*i+2] = (int32_t)( best & 0xffffffffu ); out[4*i+3] = (int32_t)( best >> 32 );
    }
  } );
}

// ---------------------------------------------------------------------------------------------------------
// Forward transform via TrQuant::xT (TrQuant.cpp:481-564).  trHor/trVer: 0 DCT2, 1 DCT8, 2 DST7.
// Output: TCoeff[h][w] row-major.  Returns 0, or -1 for a combination the explicit-MTS syntax cannot express.
int refshim_fwd_transform( int trHor, int trVer, const int16_t* resi, int stride, int w, int h, int bitDepth, int32_t* coef )
{
  RefCtx& c = ctx();
  const int mts = mtsIdxFor( trHor, trVer );
  if( mts < 0 ) return -1;
  TuRig& r = rig();
  r.setup( w, h, bitDepth, mts, false, true, 32 );
  CPelBuf  resiBuf( resi, stride, w, h );
  CoeffBuf dst( coef, w, w, h );
  tqOfThread().xT( r.tu, COMP_Y, resiBuf, dst, w, h );
  return 0;
}

// Bare 1-D core as the reference unit test exercises it (vvenc_unit_test.cpp:1085-1140): g_tCoeffOps.fastFwdCore_2D[log2N-2].
void refshim_fwd_core( int trSize, const int16_t* tc, const int32_t* src, int32_t* dst, unsigned line, unsigned reducedLine, unsigned cutoff, int shift )
{
  ctx();
  g_tCoeffOps.fastFwdCore_2D[ ilog2( trSize ) - 2 ]( tc, src, dst, line, reducedLine, cutoff, shift );
}

// Copies the reference transform matrix (CommonLib/RomTr.cpp:364-441) into out[N*N]; type 0 DCT2, 1 DCT8, 2 DST7.
int refshim_tr_matrix( int type, int N, int16_t* out )
{
  const TMatrixCoeff* p = nullptr;
  if( type == 0 ) { switch( N ) { case 2: p = g_trCoreDCT2P2[0][0]; break; case 4: p = g_trCoreDCT2P4[0][0]; break; case 8: p = g_trCoreDCT2P8[0][0]; break;
                                  case 16: p = g_trCoreDCT2P16[0][0]; break; case 32: p = g_trCoreDCT2P32[0][0]; break; case 64: p = g_trCoreDCT2P64[0][0]; break; } }
  if( type == 1 ) { switch( N ) { case 4: p = g_trCoreDCT8P4[0][0]; break; case 8: p = g_trCoreDCT8P8[0][0]; break; case 16: p = g_trCoreDCT8P16[0][0]; break; case 32: p = g_trCoreDCT8P32[0][0]; break; } }
  if( type == 2 ) { switch( N ) { case 4: p = g_trCoreDST7P4[0][0]; break; case 8: p = g_trCoreDST7P8[0][0]; break; case 16: p = g_trCoreDST7P16[0][0]; break; case 32: p = g_trCoreDST7P32[0][0]; break; } }
  if( !p ) return -1;
  memcpy( out, p, sizeof( int16_t ) * N * N );
  return 0;
}

// Scan order of the quantiser (ContextModelling.cpp:76 -> getScanOrder(SCAN_GROUPED_4x4,...), Rom.cpp:1620): raster index per scan position.
int refshim_scan_order( int w, int h, int32_t* idx )
{
  const ScanElement* s = getScanOrder( SCAN_GROUPED_4x4, ilog2( w ), ilog2( h ) );
  const int n = std::min( 32, w ) * std::min( 32, h );
  for( int i = 0; i < n; i++ ) idx[i] = s[i].idx;
  return n;
}

// Plain quantiser: Quant::quant (Quant.cpp:735-833) -> xQuant pointer (QuantCore / QuantCoreSIMD) + last-position trim.
// coef: TCoeff[h][w]; out q: int16[h][w]; returns 0.  Sign-bit hiding off (slice.signDataHidingEnabled = false).
int refshim_quant( const int32_t* coef, int w, int h, int bitDepth, int qp, int isIRAP, int intraCu, int16_t* q, int32_t* absSum, int32_t* lastPos )
{
  RefCtx& c = ctx();
  TuRig& r = rig();
  r.setup( w, h, bitDepth, MTS_DCT2_DCT2, isIRAP != 0, intraCu != 0, qp );
  QpParam qpp( r.tu, COMP_Y, false );
  CCoeffBuf src( coef, w, w, h );
  TCoeff sum = 0;
  alignas(64) static thread_local unsigned char ctxMem[ sizeof( Ctx ) ];   // unused by the plain quantiser
  const Ctx& dummy = *reinterpret_cast<const Ctx*>( ctxMem );
  static_cast<Quant*>( tqOfThread().m_quant )->Quant::quant( r.tu, COMP_Y, src, sum, qpp, dummy );
  memcpy( q, r.qcoef.data(), sizeof( int16_t ) * w * h );
  *absSum  = sum;
  *lastPos = r.tu.lastPos[COMP_Y];
  return 0;
}

// Quant::xNeedRDOQ (Quant.cpp:835-891) -> xNeedRdoq pointer (needRdoqCore / NeedRdoqSIMD)
int refshim_need_rdoq( const int32_t* coef, int w, int h, int bitDepth, int qp, int depQuant )
{
  RefCtx& c = ctx();
  TuRig& r = rig();
  r.setup( w, h, bitDepth, MTS_DCT2_DCT2, false, false, qp );
  r.slice.depQuantEnabled = depQuant != 0;
  QpParam qpp( r.tu, COMP_Y, false );
  CCoeffBuf src( coef, w, w, h );
  return static_cast<Quant*>( tqOfThread().m_quant )->xNeedRDOQ( r.tu, COMP_Y, src, qpp ) ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------------
// Transform skip and chroma components.  comp = 0 luma, 1 Cb: the rig is then built in 4:4:4 (three blocks of the TU's size) and the QpParam of the luma
// component stands in for the chroma one (its QP is whatever the caller passes -- the chroma QP mapping is host work on both sides), so that Quant::quant /
// xNeedRDOQ / dequant run with compID = COMP_Cb: the only arithmetic that differs is xNeedRDOQ's rounding constant (Quant.cpp:877).
// transformSkip: tu.mtsIdx = MTS_SKIP -> xTransformSkip (TrQuant.cpp:1050), cQP.per / rem( true ) with sps.internalMinusInputBitDepth = inputDelta.
int refshim_transform_quant_ts( const int16_t* resi, int stride, int w, int h, int bitDepth, int qp, int isIRAP, int signHiding, int depQuant, int transformSkip, int inputDelta, int comp,
                                int32_t* coef, int16_t* q, int32_t* absSum, int32_t* lastPos, int32_t* needRdoq )
{
  RefCtx& c = ctx();
  TuRig& r = rig();
  t_rigSignHiding = signHiding != 0;
  r.setup( w, h, bitDepth, MTS_DCT2_DCT2, isIRAP != 0, true, qp, comp ? CHROMA_444 : CHROMA_400 );
  t_rigSignHiding = false;
  r.slice.depQuantEnabled = depQuant != 0;
  r.sps.internalMinusInputBitDepth[CH_L] = inputDelta; r.sps.internalMinusInputBitDepth[CH_C] = inputDelta;
  const ComponentID compID = comp ? COMP_Cb : COMP_Y;
  r.tu.mtsIdx[COMP_Y] = transformSkip ? MTS_SKIP : MTS_DCT2_DCT2;          // the QpParam below is built for luma
  r.tu.mtsIdx[compID] = transformSkip ? MTS_SKIP : MTS_DCT2_DCT2;
  QpParam qpp( r.tu, COMP_Y, false );
  TrQuant& tq = tqOfThread();
  CPelBuf resiBuf( resi, stride, w, h );
  if( transformSkip ) tq.xTransformSkip( r.tu, compID, resiBuf, coef );
  else { CoeffBuf dst( coef, w, w, h ); tq.xT( r.tu, compID, resiBuf, dst, w, h ); }
  CCoeffBuf src( coef, w, w, h );
  TCoeff sum = 0;
  alignas(64) static thread_local unsigned char ctxMem[ sizeof( Ctx ) ];
  const Ctx& dummy = *reinterpret_cast<const Ctx*>( ctxMem );
  Quant* qu = static_cast<Quant*>( tq.m_quant );
  qu->Quant::quant( r.tu, compID, src, sum, qpp, dummy );
  memcpy( q, r.qcoef.data(), sizeof( int16_t ) * w * h );
  *absSum = sum; *lastPos = r.tu.lastPos[compID];
  if( needRdoq ) *needRdoq = qu->xNeedRDOQ( r.tu, compID, src, qpp ) ? 1 : 0;
  r.sps.internalMinusInputBitDepth[CH_L] = 0; r.sps.internalMinusInputBitDepth[CH_C] = 0; r.slice.depQuantEnabled = false;
  return 0;
}
// the same TU through integration/TrQuantB200.h (xTQuantB200 with compID / MTS_SKIP: vvb_tu_par.transform_skip, input_bit_depth_delta, is_chroma); 1 = the binding threw
int refshim_transform_quant_ts_b200( const int16_t* resi, int stride, int w, int h, int bitDepth, int qp, int isIRAP, int signHiding, int depQuant, int transformSkip, int inputDelta, int comp,
                                     int32_t* coef, int16_t* q, int32_t* absSum, int32_t* lastPos, int32_t* needRdoq )
{
  RefCtx& c = ctx();
  TuRig& r = rig();
  t_rigSignHiding = signHiding != 0;
  r.setup( w, h, bitDepth, MTS_DCT2_DCT2, isIRAP != 0, true, qp, comp ? CHROMA_444 : CHROMA_400 );
  t_rigSignHiding = false;
  r.slice.depQuantEnabled = depQuant != 0;
  r.sps.internalMinusInputBitDepth[CH_L] = inputDelta; r.sps.internalMinusInputBitDepth[CH_C] = inputDelta;
  const ComponentID compID = comp ? COMP_Cb : COMP_Y;
  r.tu.mtsIdx[COMP_Y] = transformSkip ? MTS_SKIP : MTS_DCT2_DCT2;
  r.tu.mtsIdx[compID] = transformSkip ? MTS_SKIP : MTS_DCT2_DCT2;
  QpParam qpp( r.tu, COMP_Y, false );
  CPelBuf resiBuf( resi, stride, w, h );
  CoeffBuf dst( coef, w, w, h );
  TCoeff sum = 0; bool nr = false;
  int rc = 0;
  try { xTQuantB200( tqOfThread(), r.tu, compID, resiBuf, dst, qpp, sum, &nr ); }
  catch( std::exception& e ) { g_b200.error = e.what(); rc = 1; }
  r.sps.internalMinusInputBitDepth[CH_L] = 0; r.sps.internalMinusInputBitDepth[CH_C] = 0; r.slice.depQuantEnabled = false;
  if( rc ) return rc;
  memcpy( q, r.qcoef.data(), sizeof( int16_t ) * w * h );
  *absSum = sum; *lastPos = r.tu.lastPos[compID]; *needRdoq = nr ? 1 : 0;
  return 0;
}
int refshim_inv_transform_quant_ts_b200( const int16_t* q, int w, int h, int bitDepth, int qp, int inputDelta, int16_t* resi, int stride )
{
  RefCtx& c = ctx();
  TuRig& r = rig();
  r.setup( w, h, bitDepth, MTS_SKIP, false, true, qp );
  r.sps.internalMinusInputBitDepth[CH_L] = inputDelta;
  memcpy( r.qcoef.data(), q, sizeof( int16_t ) * w * h );
  QpParam qpp( r.tu, COMP_Y, false );
  PelBuf out( resi, stride, w, h );
  int rc = 0;
  try { invTransformNxNB200( tqOfThread(), r.tu, COMP_Y, out, qpp ); }
  catch( std::exception& e ) { g_b200.error = e.what(); rc = 1; }
  r.sps.internalMinusInputBitDepth[CH_L] = 0;
  return rc;
}

// TrQuant::invTransformNxN for a skipped transform: Quant::dequant + xITransformSkip
int refshim_inv_transform_quant_ts( const int16_t* q, int w, int h, int bitDepth, int qp, int inputDelta, int32_t* coef, int16_t* resi, int stride )
{
  RefCtx& c = ctx();
  TuRig& r = rig();
  r.setup( w, h, bitDepth, MTS_SKIP, false, true, qp );
  r.sps.internalMinusInputBitDepth[CH_L] = inputDelta;
  memcpy( r.qcoef.data(), q, sizeof( int16_t ) * w * h );
  QpParam qpp( r.tu, COMP_Y, false );
  alignas(64) static thread_local TCoeff tmpCoef[ 64 * 64 ];
  CoeffBuf deq( tmpCoef, w, w, h );
  static_cast<Quant*>( tqOfThread().m_quant )->Quant::dequant( r.tu, deq, COMP_Y, qpp );
  if( coef ) memcpy( coef, tmpCoef, sizeof( int32_t ) * w * h );
  PelBuf out( resi, stride, w, h );
  tqOfThread().xITransformSkip( CCoeffBuf( tmpCoef, w, w, h ), out, r.tu, COMP_Y );
  r.sps.internalMinusInputBitDepth[CH_L] = 0;
  return 0;
}

// Whole TU: xT then Quant::quant, as TrQuant::transformNxN does for LFNST-off, non-skip TUs (TrQuant.cpp:688-736).
int refshim_transform_quant( int trHor, int trVer, const int16_t* resi, int stride, int w, int h, int bitDepth, int qp, int isIRAP,
                             int32_t* coef, int16_t* q, int32_t* absSum, int32_t* lastPos )
{
  RefCtx& c = ctx();
  const int mts = mtsIdxFor( trHor, trVer );
  if( mts < 0 ) return -1;
  TuRig& r = rig();
  r.setup( w, h, bitDepth, mts, isIRAP != 0, true, qp );
  CPelBuf  resiBuf( resi, stride, w, h );
  CoeffBuf dst( coef, w, w, h );
  tqOfThread().xT( r.tu, COMP_Y, resiBuf, dst, w, h );
  QpParam qpp( r.tu, COMP_Y, false );
  CCoeffBuf src( coef, w, w, h );
  TCoeff sum = 0;
  alignas(64) static thread_local unsigned char ctxMem[ sizeof( Ctx ) ];
  const Ctx& dummy = *reinterpret_cast<const Ctx*>( ctxMem );
  static_cast<Quant*>( tqOfThread().m_quant )->Quant::quant( r.tu, COMP_Y, src, sum, qpp, dummy );
  memcpy( q, r.qcoef.data(), sizeof( int16_t ) * w * h );
  *absSum = sum; *lastPos = r.tu.lastPos[COMP_Y];
  return 0;
}

// same with slice->signDataHidingEnabled selectable: Quant::quant then runs xSignBitHidingHDQ (Quant.cpp:817-826, 377-518)
int refshim_transform_quant_sdh( int trHor, int trVer, const int16_t* resi, int stride, int w, int h, int bitDepth, int qp, int isIRAP, int signHiding,
                                 int32_t* coef, int16_t* q, int32_t* absSum, int32_t* lastPos )
{
  t_rigSignHiding = signHiding != 0;
  const int rc = refshim_transform_quant( trHor, trVer, resi, stride, w, h, bitDepth, qp, isIRAP, coef, q, absSum, lastPos );
  t_rigSignHiding = false;
  return rc;
}


// ---------------------------------------------------------------------------------------------------------
// LFNST: TrQuant::transformNxN's own sequence for a luma TU of an intra CU with cu.lfnstIdx = lfnstIdx -- xT (zero-out of :499-511), xFwdLfnst (:942-1048,
// through m_fwdLfnstNxN), Quant::quant -- on the TU rig.  xFwdLfnst looks the CU up through the coding structure, so the rig's CodingStructure gets a
// one-CU map for the call.  outSetTranspose[0] = g_lfnstLut[ xGetLFNSTIntraMode( intraMode ) ], [1] = xGetTransposeFlag: what the B200 binding passes down.
int refshim_transform_quant_lfnst( const int16_t* resi, int stride, int w, int h, int bitDepth, int qp, int isIRAP, int signHiding, int intraMode, int lfnstIdx,
                                   int32_t* coef, int16_t* q, int32_t* absSum, int32_t* lastPos, int32_t* needRdoq, int32_t* outSetTranspose )
{
  RefCtx& c = ctx();
  TuRig& r = rig();
  t_rigSignHiding = signHiding != 0;
  r.setup( w, h, bitDepth, 0, isIRAP != 0, true, qp );
  t_rigSignHiding = false;
  r.sps.LFNST = true;
  r.slice.sliceType = VVENC_B_SLICE;          // CU::isSepTree would consult cs.pcv (absent on the rig) for an I slice; the rounding offset follows isIRAP() = the NAL unit type
  r.cu.lfnstIdx = (uint8_t) lfnstIdx;
  r.cu.intraDir[CH_L] = (uint8_t) intraMode; r.cu.intraDir[CH_C] = DM_CHROMA_IDX;
  r.cu.mipFlag = false; r.cu.ispMode = 0; r.cu.chromaFormat = CHROMA_400;
  // one-CU lookup map for CodingStructure::getCU (CodingStructure.cpp:140-156)
  static thread_local std::vector<CodingUnit*> cuMap;
  r.cs.area = UnitArea( CHROMA_400, Area( 0, 0, w, h ) );
  r.cs.parent = nullptr;
  r.cs.unitScale[COMP_Y] = UnitScale( MIN_CU_LOG2, MIN_CU_LOG2 );
  cuMap.assign( (size_t)( ( w >> MIN_CU_LOG2 ) + 1 ) * ( ( h >> MIN_CU_LOG2 ) + 1 ), &r.cu );
  r.cs.m_cuPtr[CH_L] = cuMap.data();
  TrQuant& tq = tqOfThread();
  CPelBuf  resiBuf( resi, stride, w, h );
  CoeffBuf tmp( tq.m_plTempCoeff, w, w, h );
  tq.xT( r.tu, COMP_Y, resiBuf, tmp, w, h );
  tq.xFwdLfnst( r.tu, COMP_Y, false );
  for( int y = 0; y < h; y++ ) memcpy( coef + (size_t) y * w, tq.m_plTempCoeff + (size_t) y * w, sizeof( TCoeff ) * w );
  QpParam qpp( r.tu, COMP_Y, false );
  CCoeffBuf src( coef, w, w, h );
  TCoeff sum = 0;
  alignas(64) static thread_local unsigned char ctxMem[ sizeof( Ctx ) ];
  const Ctx& dummy = *reinterpret_cast<const Ctx*>( ctxMem );
  Quant* qu = static_cast<Quant*>( tq.m_quant );
  qu->Quant::quant( r.tu, COMP_Y, src, sum, qpp, dummy );
  memcpy( q, r.qcoef.data(), sizeof( int16_t ) * w * h );
  *absSum = sum; *lastPos = r.tu.lastPos[COMP_Y];
  if( needRdoq ) *needRdoq = qu->xNeedRDOQ( r.tu, COMP_Y, src, qpp ) ? 1 : 0;
  const uint32_t m = tq.xGetLFNSTIntraMode( r.tu.blocks[COMP_Y], (uint32_t) intraMode );
  outSetTranspose[0] = g_lfnstLut[m]; outSetTranspose[1] = tq.xGetTransposeFlag( m ) ? 1 : 0;
  r.cs.m_cuPtr[CH_L] = nullptr;
  r.cu.lfnstIdx = 0; r.sps.LFNST = false;
  return 0;
}
// TrQuant::invTransformNxN of the same kind of TU (the member itself: xDeQuant, xInvLfnst :838-940, xIT with the LFNST skip :590-602).  coefOut: the buffer xInvLfnst leaves
int refshim_inv_transform_quant_lfnst( const int16_t* q, int w, int h, int bitDepth, int qp, int depQuant, int lastPos, int intraMode, int lfnstIdx, int32_t* coefOut, int16_t* resi, int stride,
                                       int32_t* outSetTranspose )
{
  RefCtx& c = ctx();
  TuRig& r = rig();
  r.setup( w, h, bitDepth, 0, false, true, qp );
  r.slice.depQuantEnabled = depQuant != 0;
  r.tu.lastPos[COMP_Y] = lastPos;
  r.sps.LFNST = true;
  r.slice.sliceType = VVENC_B_SLICE;
  r.cu.lfnstIdx = (uint8_t) lfnstIdx;
  r.cu.intraDir[CH_L] = (uint8_t) intraMode; r.cu.intraDir[CH_C] = DM_CHROMA_IDX;
  r.cu.mipFlag = false; r.cu.ispMode = 0; r.cu.chromaFormat = CHROMA_400;
  static thread_local std::vector<CodingUnit*> cuMap;
  r.cs.area = UnitArea( CHROMA_400, Area( 0, 0, w, h ) );
  r.cs.parent = nullptr;
  r.cs.unitScale[COMP_Y] = UnitScale( MIN_CU_LOG2, MIN_CU_LOG2 );
  cuMap.assign( (size_t)( ( w >> MIN_CU_LOG2 ) + 1 ) * ( ( h >> MIN_CU_LOG2 ) + 1 ), &r.cu );
  r.cs.m_cuPtr[CH_L] = cuMap.data();
  memcpy( r.qcoef.data(), q, sizeof( int16_t ) * w * h );
  TrQuant& tq = tqOfThread();
  QpParam qpp( r.tu, COMP_Y, false );
  PelBuf out( resi, stride, w, h );
  tq.invTransformNxN( r.tu, COMP_Y, out, qpp );
  if( coefOut ) for( int y = 0; y < h; y++ ) memcpy( coefOut + (size_t) y * w, tq.m_plTempCoeff + (size_t) y * w, sizeof( TCoeff ) * w );
  const uint32_t m = tq.xGetLFNSTIntraMode( r.tu.blocks[COMP_Y], (uint32_t) intraMode );
  outSetTranspose[0] = g_lfnstLut[m]; outSetTranspose[1] = tq.xGetTransposeFlag( m ) ? 1 : 0;
  r.cs.m_cuPtr[CH_L] = nullptr;
  r.cu.lfnstIdx = 0; r.sps.LFNST = false; r.slice.depQuantEnabled = false;
  return 0;
}
// the inverse through integration/TrQuantB200.h (invTransformNxNB200 derives kernel set / transposition from the CU); returns 1 when the binding threw
int refshim_inv_transform_quant_lfnst_b200( const int16_t* q, int w, int h, int bitDepth, int qp, int depQuant, int lastPos, int intraMode, int lfnstIdx, int16_t* resi, int stride )
{
  RefCtx& c = ctx();
  TuRig& r = rig();
  r.setup( w, h, bitDepth, 0, false, true, qp );
  r.slice.depQuantEnabled = depQuant != 0;
  r.tu.lastPos[COMP_Y] = lastPos;
  r.sps.LFNST = true;
  r.slice.sliceType = VVENC_B_SLICE;
  r.cu.lfnstIdx = (uint8_t) lfnstIdx;
  r.cu.intraDir[CH_L] = (uint8_t) intraMode; r.cu.intraDir[CH_C] = DM_CHROMA_IDX;
  r.cu.mipFlag = false; r.cu.ispMode = 0; r.cu.chromaFormat = CHROMA_400;
  memcpy( r.qcoef.data(), q, sizeof( int16_t ) * w * h );
  QpParam qpp( r.tu, COMP_Y, false );
  PelBuf out( resi, stride, w, h );
  int rc = 0;
  try { invTransformNxNB200( tqOfThread(), r.tu, COMP_Y, out, qpp ); }
  catch( std::exception& e ) { g_b200.error = e.what(); rc = 1; }
  r.cu.lfnstIdx = 0; r.sps.LFNST = false; r.slice.depQuantEnabled = false;
  return rc;
}
// the same TU through integration/TrQuantB200.h (xTQuantB200 derives kernel set / transposition from the CU like xFwdLfnst does); returns 1 when the binding threw
int refshim_transform_quant_lfnst_b200( const int16_t* resi, int stride, int w, int h, int bitDepth, int qp, int isIRAP, int signHiding, int intraMode, int lfnstIdx,
                                        int32_t* coef, int16_t* q, int32_t* absSum, int32_t* lastPos, int32_t* needRdoq )
{
  RefCtx& c = ctx();
  TuRig& r = rig();
  t_rigSignHiding = signHiding != 0;
  r.setup( w, h, bitDepth, 0, isIRAP != 0, true, qp );
  t_rigSignHiding = false;
  r.sps.LFNST = true;
  r.slice.sliceType = VVENC_B_SLICE;
  r.cu.lfnstIdx = (uint8_t) lfnstIdx;
  r.cu.intraDir[CH_L] = (uint8_t) intraMode; r.cu.intraDir[CH_C] = DM_CHROMA_IDX;
  r.cu.mipFlag = false; r.cu.ispMode = 0; r.cu.chromaFormat = CHROMA_400;
  CPelBuf  resiBuf( resi, stride, w, h );
  CoeffBuf dst( coef, w, w, h );
  QpParam qpp( r.tu, COMP_Y, false );
  TCoeff sum = 0; bool nr = false;
  int rc = 0;
  try { xTQuantB200( tqOfThread(), r.tu, COMP_Y, resiBuf, dst, qpp, sum, &nr ); }
  catch( std::exception& e ) { g_b200.error = e.what(); rc = 1; }
  r.cu.lfnstIdx = 0; r.sps.LFNST = false;
  if( rc ) return rc;
  memcpy( q, r.qcoef.data(), sizeof( int16_t ) * w * h );
  *absSum = sum; *lastPos = r.tu.lastPos[COMP_Y]; *needRdoq = nr ? 1 : 0;
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Dependent quantisation: DepQuant::xQuantDQ (DepQuant.cpp:1129-1264) -- what DepQuant::quant (:1462-1490) calls for a non-skip TU of a slice with
// depQuantEnabled, scaling lists off -- on the TU rig, with a CABAC context set initialised the way a slice start does it (Ctx::init( qp, initId )).
// enableOpt selects the members the DepQuant constructor installs: 0 = scalar (DQIntern::checkAllRdCosts, updateStates, ...), 1 = initDepQuantX86's.
// ratesOut: the RateEstimator tables initCtx left, in vvb_dq_rates layout (266 int32); quantOut: the 9 Quantizer constants.
int refshim_dep_quant_comp( int comp, const int32_t* coef, int w, int h, int bitDepth, int qp, int mtsIdx, int intraCu, int lfnstIdx, int sbtInfo, double lambda, int dqThrVal, int enableOpt,
                            int ctxQp, int ctxInitId, int16_t* q, int32_t* absSum, int32_t* lastPos, int32_t* ratesOut, int64_t* quantOut );
int refshim_dep_quant( const int32_t* coef, int w, int h, int bitDepth, int qp, int mtsIdx, int intraCu, int lfnstIdx, int sbtInfo, double lambda, int dqThrVal, int enableOpt,
                       int ctxQp, int ctxInitId, int16_t* q, int32_t* absSum, int32_t* lastPos, int32_t* ratesOut, int64_t* quantOut )
{
  return refshim_dep_quant_comp( 0, coef, w, h, bitDepth, qp, mtsIdx, intraCu, lfnstIdx, sbtInfo, lambda, dqThrVal, enableOpt, ctxQp, ctxInitId, q, absSum, lastPos, ratesOut, quantOut );
}
// comp = 1: the Cb component of a 4:4:4 rig (the QpParam of the luma component stands in: the chroma QP mapping is host work on both sides)
int refshim_dep_quant_comp( int comp, const int32_t* coef, int w, int h, int bitDepth, int qp, int mtsIdx, int intraCu, int lfnstIdx, int sbtInfo, double lambda, int dqThrVal, int enableOpt,
                            int ctxQp, int ctxInitId, int16_t* q, int32_t* absSum, int32_t* lastPos, int32_t* ratesOut, int64_t* quantOut )
{
  RefCtx& c = ctx();
  TuRig& r = rig();
  const ComponentID compID = comp ? COMP_Cb : COMP_Y;
  r.setup( w, h, bitDepth, mtsIdx, false, intraCu != 0, qp, comp ? CHROMA_444 : CHROMA_400 );
  if( comp ) r.tu.mtsIdx[COMP_Cb] = 0;
  r.slice.depQuantEnabled = true;
  r.cu.lfnstIdx = (uint8_t) lfnstIdx; r.cu.sbtInfo = (uint8_t) sbtInfo;
  static thread_local std::unique_ptr<DepQuant> dqs[2];
  std::unique_ptr<DepQuant>& dq = dqs[enableOpt ? 1 : 0];
  if( !dq ) dq.reset( new DepQuant( nullptr, true, false, enableOpt != 0 ) );
  dq->init( 0, false, dqThrVal );
  static thread_local std::unique_ptr<Ctx> cabac;
  if( !cabac ) cabac.reset( new Ctx( (const BinProbModel*) nullptr ) );
  cabac->init( ctxQp, ctxInitId );
  QpParam qpp( r.tu, COMP_Y, false );
  CCoeffBuf src( coef, w, w, h );
  TCoeff sum = 0;
  dq->xQuantDQ( r.tu, src, compID, qpp, lambda, *cabac, sum, false, nullptr );
  memcpy( q, r.qcoef.data(), sizeof( int16_t ) * w * h );
  *absSum = sum; *lastPos = r.tu.lastPos[compID];
  if( ratesOut )
  {
    // the estimator is only initialised when a first position was found (:1192-1199): run it here in any case so that callers always get the tables
    const DQIntern::TUParameters& tuPars = *dq->m_scansRom->getTUPars( r.tu.blocks[compID], compID );
    ( (DQIntern::RateEstimator&) *dq ).initCtx( tuPars, r.tu, compID, cabac->getFracBitsAcess() );
    const DQIntern::RateEstimator& re = (const DQIntern::RateEstimator&) *dq;      // private base: only a C-style cast reaches it
    int32_t* o = ratesOut;
    for( int i = 0; i < 32; i++ ) *o++ = re.m_lastBitsX[i];
    for( int i = 0; i < 32; i++ ) *o++ = re.m_lastBitsY[i];
    for( int i = 0; i < 2; i++ ) for( int b = 0; b < 2; b++ ) *o++ = re.m_sigSbbFracBits[i].intBits[b];
    for( int s = 0; s < 3; s++ ) for( int i = 0; i < 12; i++ ) for( int b = 0; b < 2; b++ ) *o++ = re.m_sigFracBits[s][i].intBits[b];
    for( int i = 0; i < 21; i++ ) for( int b = 0; b < 6; b++ ) *o++ = re.m_gtxFracBits[i].bits[b];
  }
  if( quantOut )
  {
    dq->m_quant.initQuantBlock( r.tu, compID, qpp, lambda );
    const DQIntern::Quantizer& z = dq->m_quant;
    quantOut[0] = z.m_QShift; quantOut[1] = z.m_maxQIdx; quantOut[2] = z.m_thresLast; quantOut[3] = z.m_DistShift; quantOut[4] = z.m_QAdd; quantOut[5] = z.m_QScale;
    quantOut[6] = z.m_DistAdd; quantOut[7] = z.m_DistStepAdd; quantOut[8] = z.m_DistOrgFact;
  }
  r.cu.lfnstIdx = 0; r.cu.sbtInfo = 0; r.slice.depQuantEnabled = false;
  return 0;
}

// the same TU through integration/TrQuantB200.h (xQuantDQB200: rate tables from the CABAC state here, trellis in the bound library); returns 1 when the binding threw
int refshim_dep_quant_b200_comp( int comp, const int32_t* coef, int w, int h, int bitDepth, int qp, int mtsIdx, int intraCu, int lfnstIdx, int sbtInfo, double lambda, int dqThrVal,
                                 int ctxQp, int ctxInitId, int16_t* q, int32_t* absSum, int32_t* lastPos );
int refshim_dep_quant_b200( const int32_t* coef, int w, int h, int bitDepth, int qp, int mtsIdx, int intraCu, int lfnstIdx, int sbtInfo, double lambda, int dqThrVal,
                            int ctxQp, int ctxInitId, int16_t* q, int32_t* absSum, int32_t* lastPos )
{
  return refshim_dep_quant_b200_comp( 0, coef, w, h, bitDepth, qp, mtsIdx, intraCu, lfnstIdx, sbtInfo, lambda, dqThrVal, ctxQp, ctxInitId, q, absSum, lastPos );
}
int refshim_dep_quant_b200_comp( int comp, const int32_t* coef, int w, int h, int bitDepth, int qp, int mtsIdx, int intraCu, int lfnstIdx, int sbtInfo, double lambda, int dqThrVal,
                                 int ctxQp, int ctxInitId, int16_t* q, int32_t* absSum, int32_t* lastPos )
{
  RefCtx& c = ctx();
  TuRig& r = rig();
  const ComponentID compID = comp ? COMP_Cb : COMP_Y;
  r.setup( w, h, bitDepth, mtsIdx, false, intraCu != 0, qp, comp ? CHROMA_444 : CHROMA_400 );
  if( comp ) r.tu.mtsIdx[COMP_Cb] = 0;
  r.slice.depQuantEnabled = true;
  r.cu.lfnstIdx = (uint8_t) lfnstIdx; r.cu.sbtInfo = (uint8_t) sbtInfo;
  if( lfnstIdx ) { r.sps.LFNST = true; r.cu.intraDir[CH_L] = PLANAR_IDX; r.cu.intraDir[CH_C] = DM_CHROMA_IDX; r.cu.mipFlag = false; r.cu.ispMode = 0; r.cu.chromaFormat = CHROMA_400; }
  static thread_local std::unique_ptr<DepQuant> dq;
  if( !dq ) dq.reset( new DepQuant( nullptr, true, false, true ) );
  dq->init( 0, false, dqThrVal );
  static thread_local std::unique_ptr<Ctx> cabac;
  if( !cabac ) cabac.reset( new Ctx( (const BinProbModel*) nullptr ) );
  cabac->init( ctxQp, ctxInitId );
  QpParam qpp( r.tu, COMP_Y, false );
  CCoeffBuf src( coef, w, w, h );
  TCoeff sum = 0;
  int rc = 0;
  try { xQuantDQB200( *dq, tqOfThread(), r.tu, src, compID, qpp, lambda, *cabac, sum ); }
  catch( std::exception& e ) { g_b200.error = e.what(); rc = 1; }
  r.cu.lfnstIdx = 0; r.cu.sbtInfo = 0; r.slice.depQuantEnabled = false; r.sps.LFNST = false;
  if( rc ) return rc;
  memcpy( q, r.qcoef.data(), sizeof( int16_t ) * w * h );
  *absSum = sum; *lastPos = r.tu.lastPos[compID];
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Fast RDOQ: QuantRDOQ2::xRateDistOptQuant (QuantRDOQ2.cpp:1283-1296 -> xRateDistOptQuantFast<bSBH, false> :475-1281) -- what QuantRDOQ2::quant (:247-301) runs with
// m_RDOQ == 2 for a TU that is not transform skipped -- on the TU rig, with a CABAC context set initialised the way a slice start does it.  The object is a DepQuant
// (the class TrQuant instantiates; QuantRDOQ2 is its base) after init( 2, ., thrVal ) and setFlatScalingList (the error scales, EncCu.cpp:351).
// ratesOut: the fractional bits the routine read, in vvb_rdoq_rates layout (190 int32); constOut: quantScale, errScale, qBits, useThres, remRegBins, numCG, firstScanPos.
// cbCbf: tu.cbf[COMP_Cb] when comp == 2 (Cr reuses the last-position table of the Cb call, :490, and its coded-block-flag context depends on it).
static void rdoqRatesOf( DepQuant& dq, const TransformUnit& tu, const ComponentID compID, const Ctx& cabac, int32_t* o )
{
  const ChannelType ch = toChannelType( compID );
  const FracBitsAccess& fb = cabac.getFracBitsAcess();
  memset( o, 0, sizeof( int32_t ) * 190 );
  int32_t* sig = o, *par = o + 24, *gt1 = par + 42, *gt2 = gt1 + 42, *grp = gt2 + 42, *lx = grp + 4, *ly = lx + 16, *cbf = ly + 16;
  for( int i = 0; i < (int) Ctx::SigFlag[ch].size() && i < 12; i++ ) for( int b = 0; b < 2; b++ ) sig[2 * i + b] = fb.getFracBitsArray( Ctx::SigFlag[ch]( i ) ).intBits[b];
  for( int i = 0; i < (int) Ctx::ParFlag[ch].size() && i < 21; i++ ) for( int b = 0; b < 2; b++ ) par[2 * i + b] = fb.getFracBitsArray( Ctx::ParFlag[ch]( i ) ).intBits[b];
  for( int i = 0; i < (int) Ctx::GtxFlag[ch + 2].size() && i < 21; i++ ) for( int b = 0; b < 2; b++ ) gt1[2 * i + b] = fb.getFracBitsArray( Ctx::GtxFlag[ch + 2]( i ) ).intBits[b];
  for( int i = 0; i < (int) Ctx::GtxFlag[ch].size() && i < 21; i++ ) for( int b = 0; b < 2; b++ ) gt2[2 * i + b] = fb.getFracBitsArray( Ctx::GtxFlag[ch]( i ) ).intBits[b];
  for( int i = 0; i < 2; i++ ) for( int b = 0; b < 2; b++ ) grp[2 * i + b] = fb.getFracBitsArray( Ctx::SigCoeffGroup[ch]( i ) ).intBits[b];
  for( int i = 0; i < LAST_SIGNIFICANT_GROUPS; i++ ) { lx[i] = dq.QuantRDOQ2::m_lastBitsX[ch][i]; ly[i] = dq.QuantRDOQ2::m_lastBitsY[ch][i]; }
  if( !CU::isIntra( *tu.cu ) && isLuma( compID ) ) { const BinFracBits f = fb.getFracBitsArray( Ctx::QtRootCbf() ); cbf[0] = f.intBits[0]; cbf[1] = f.intBits[1]; }
  else { const BinFracBits f = fb.getFracBitsArray( Ctx::QtCbf[compID]( DeriveCtx::CtxQtCbf( compID, tu.cbf[COMP_Cb], false ) ) ); cbf[0] = f.intBits[0]; cbf[1] = f.intBits[1]; }
}
int refshim_rdoq( int comp, const int32_t* coef, int w, int h, int bitDepth, int qp, int intraCu, int lfnstIdx, int sbtInfo, int signHiding, int cbCbf, double lambda, int thrVal,
                  int ctxQp, int ctxInitId, int16_t* q, int32_t* absSum, int32_t* lastPos, int32_t* ratesOut, int32_t* constOut )
{
  RefCtx& c = ctx(); (void) c;
  TuRig& r = rig();
  const ComponentID compID = comp == 2 ? COMP_Cr : comp ? COMP_Cb : COMP_Y;
  t_rigSignHiding = signHiding != 0;
  r.setup( w, h, bitDepth, 0, false, intraCu != 0, qp, comp ? CHROMA_444 : CHROMA_400 );
  t_rigSignHiding = false;
  r.cu.lfnstIdx = (uint8_t) lfnstIdx; r.cu.sbtInfo = (uint8_t) sbtInfo;
  r.tu.cbf[COMP_Cb] = (uint8_t)( cbCbf ? 1 : 0 );
  static thread_local std::unique_ptr<DepQuant> dq;
  if( !dq ) dq.reset( new DepQuant( nullptr, true, false, true ) );
  dq->init( 2, false, thrVal );
  const int maxLog2TrDynamicRange[MAX_NUM_CH] = { 15, 15 };
  dq->setFlatScalingList( maxLog2TrDynamicRange, r.sps.bitDepths );
  dq->m_dLambda = lambda;
  static thread_local std::unique_ptr<Ctx> cabac;
  if( !cabac ) cabac.reset( new Ctx( (const BinProbModel*) nullptr ) );
  cabac->init( ctxQp, ctxInitId );
  QpParam qpp( r.tu, COMP_Y, false );
  CCoeffBuf src( coef, w, w, h );
  TCoeff sum = 0;
  if( comp == 2 && cbCbf )
  {
    // the Cb call that precedes a Cr call in the encoder fills the chroma last-position table (:490): run the table set-up it would have run
    CoeffCodingContext cctx( r.tu, COMP_Cb, signHiding != 0, false, dq->m_tplBuf );
    dq->xInitLastPosBitsTab( cctx, w, h, CH_C, cabac->getFracBitsAcess() );
  }
  dq->xRateDistOptQuant( r.tu, compID, src, sum, qpp, *cabac, false );
  memcpy( q, r.qcoef.data(), sizeof( int16_t ) * w * h );
  *absSum = sum; *lastPos = r.tu.lastPos[compID];
  if( ratesOut ) rdoqRatesOf( *dq, r.tu, compID, *cabac, ratesOut );
  if( constOut )
  {
    const int rem = qpp.rem( false ), per = qpp.per( false );
    const bool sqrt2 = TU::needsSqrt2Scale( r.tu, compID );
    const int trShift = getTransformShift( bitDepth, r.tu.blocks[compID].size(), 15 );
    constOut[0] = g_quantScales[sqrt2 ? 1 : 0][rem];
    constOut[1] = dq->xGetErrScaleCoeffNoScalingList( getScalingListType( r.cu.predMode, compID ), Log2( w ), Log2( h ), rem );
    constOut[2] = QUANT_SHIFT + per + trShift + ( sqrt2 ? -1 : 0 );
    const TCoeff thres = constOut[2] ? TCoeff( ( int64_t( thrVal ) << ( constOut[2] - 1 ) ) ) : TCoeff( ( int64_t( thrVal >> 1 ) << constOut[2] ) );
    constOut[3] = thres / ( constOut[0] << 2 );
    constOut[4] = ( r.tu.getTbAreaAfterCoefZeroOut( compID ) * MAX_TU_LEVEL_CTX_CODED_BIN_CONSTRAINT ) >> 4;
    constOut[5] = lfnstIdx > 0 ? 1 : ( std::min<int>( JVET_C0024_ZERO_OUT_TH, w ) * std::min<int>( JVET_C0024_ZERO_OUT_TH, h ) ) >> 4;
    constOut[6] = ( lfnstIdx > 0 && ( ( w == 4 && h == 4 ) || ( w == 8 && h == 8 ) ) ) ? 7 : ( constOut[5] << 4 ) - 1;
  }
  r.cu.lfnstIdx = 0; r.cu.sbtInfo = 0; r.tu.cbf[COMP_Cb] = 0;
  return 0;
}

// the same TU through integration/TrQuantB200.h (xRateDistOptQuantB200: rates from the CABAC state here, level decisions in the bound library); returns 1 when the binding threw
int refshim_rdoq_b200( int comp, const int32_t* coef, int w, int h, int bitDepth, int qp, int intraCu, int lfnstIdx, int sbtInfo, int signHiding, int cbCbf, double lambda, int thrVal,
                       int ctxQp, int ctxInitId, int16_t* q, int32_t* absSum, int32_t* lastPos )
{
  RefCtx& c = ctx(); (void) c;
  TuRig& r = rig();
  const ComponentID compID = comp == 2 ? COMP_Cr : comp ? COMP_Cb : COMP_Y;
  t_rigSignHiding = signHiding != 0;
  r.setup( w, h, bitDepth, 0, false, intraCu != 0, qp, comp ? CHROMA_444 : CHROMA_400 );
  t_rigSignHiding = false;
  r.cu.lfnstIdx = (uint8_t) lfnstIdx; r.cu.sbtInfo = (uint8_t) sbtInfo;
  r.tu.cbf[COMP_Cb] = (uint8_t)( cbCbf ? 1 : 0 );
  static thread_local std::unique_ptr<DepQuant> dq;
  if( !dq ) dq.reset( new DepQuant( nullptr, true, false, true ) );
  dq->init( 2, false, thrVal );
  dq->m_dLambda = lambda;
  static thread_local std::unique_ptr<Ctx> cabac;
  if( !cabac ) cabac.reset( new Ctx( (const BinProbModel*) nullptr ) );
  cabac->init( ctxQp, ctxInitId );
  QpParam qpp( r.tu, COMP_Y, false );
  CCoeffBuf src( coef, w, w, h );
  TCoeff sum = 0;
  if( comp == 2 && cbCbf )
  {
    CoeffCodingContext cctx( r.tu, COMP_Cb, signHiding != 0, false, dq->m_tplBuf );
    dq->xInitLastPosBitsTab( cctx, w, h, CH_C, cabac->getFracBitsAcess() );
  }
  int rc = 0;
  try { xRateDistOptQuantB200( *dq, tqOfThread(), r.tu, compID, src, sum, qpp, *cabac ); }
  catch( std::exception& e ) { g_b200.error = e.what(); rc = 1; }
  r.cu.lfnstIdx = 0; r.cu.sbtInfo = 0; r.tu.cbf[COMP_Cb] = 0;
  if( rc ) return rc;
  memcpy( q, r.qcoef.data(), sizeof( int16_t ) * w * h );
  *absSum = sum; *lastPos = r.tu.lastPos[compID];
  return 0;
}

// Transform-skip RDOQ: QuantRDOQ::rateDistOptQuantTS (QuantRDOQ.cpp:1124-1336) on the TU rig with tu.mtsIdx = MTS_SKIP.  ratesOut: the 44 int32 of vvb_rdoq_ts_rates;
// constOut: quantScale, qBits, maxCtxBins; errScaleOut: xGetErrScaleCoeff( ..., true ).
static void rdoqTsRatesOf( const Ctx& cabac, int32_t* o )
{
  const FracBitsAccess& fb = cabac.getFracBitsAcess();
  for( int i = 0; i < 3; i++ ) for( int b = 0; b < 2; b++ ) *o++ = fb.getFracBitsArray( Ctx::TsSigFlag( i ) ).intBits[b];
  for( int b = 0; b < 2; b++ ) *o++ = fb.getFracBitsArray( Ctx::TsParFlag( 0 ) ).intBits[b];
  for( int i = 0; i < 5; i++ ) for( int b = 0; b < 2; b++ ) *o++ = fb.getFracBitsArray( Ctx::TsGtxFlag( i ) ).intBits[b];
  for( int i = 0; i < 4; i++ ) for( int b = 0; b < 2; b++ ) *o++ = fb.getFracBitsArray( Ctx::TsLrg1Flag( i ) ).intBits[b];
  for( int i = 0; i < 6; i++ ) for( int b = 0; b < 2; b++ ) *o++ = fb.getFracBitsArray( Ctx::TsResidualSign( i ) ).intBits[b];
  for( int i = 0; i < 3; i++ ) for( int b = 0; b < 2; b++ ) *o++ = fb.getFracBitsArray( Ctx::TsSigCoeffGroup( i ) ).intBits[b];
}
int refshim_rdoq_ts( int comp, const int32_t* coef, int w, int h, int bitDepth, int qp, int inputDelta, int intraCu, double lambda, int ctxQp, int ctxInitId,
                     int16_t* q, int32_t* absSum, int32_t* ratesOut, int32_t* constOut, double* errScaleOut )
{
  RefCtx& c = ctx(); (void) c;
  TuRig& r = rig();
  const ComponentID compID = comp ? COMP_Cb : COMP_Y;
  r.setup( w, h, bitDepth, MTS_SKIP, false, intraCu != 0, qp, comp ? CHROMA_444 : CHROMA_400 );
  if( comp ) r.tu.mtsIdx[COMP_Cb] = MTS_SKIP;
  r.sps.internalMinusInputBitDepth[CH_L] = inputDelta; r.sps.internalMinusInputBitDepth[CH_C] = inputDelta;
  static thread_local std::unique_ptr<DepQuant> dq;
  if( !dq ) dq.reset( new DepQuant( nullptr, true, false, true ) );
  dq->init( 2, true, 8 );
  dq->m_dLambda = lambda;
  static thread_local std::unique_ptr<Ctx> cabac;
  if( !cabac ) cabac.reset( new Ctx( (const BinProbModel*) nullptr ) );
  cabac->init( ctxQp, ctxInitId );
  QpParam qpp( r.tu, COMP_Y, false );
  CCoeffBuf src( coef, w, w, h );
  TCoeff sum = 0;
  dq->rateDistOptQuantTS( r.tu, compID, src, sum, qpp, *cabac );
  memcpy( q, r.qcoef.data(), sizeof( int16_t ) * w * h );
  *absSum = sum;
  if( ratesOut ) rdoqTsRatesOf( *cabac, ratesOut );
  if( constOut ) { constOut[0] = g_quantScales[0][qpp.rem( true )]; constOut[1] = QUANT_SHIFT + qpp.per( true ); constOut[2] = ( w * h * 7 ) >> 2; }
  if( errScaleOut ) *errScaleOut = dq->QuantRDOQ::xGetErrScaleCoeff( false, w, h, qpp.rem( true ), 15, bitDepth, true );
  r.sps.internalMinusInputBitDepth[CH_L] = 0; r.sps.internalMinusInputBitDepth[CH_C] = 0;
  return 0;
}

// BDPCM: QuantRDOQ::forwardRDPCM (QuantRDOQ.cpp:1338-1562) on the rig with cu.bdpcmM = dirMode (1 horizontal, 2 vertical); rates as refshim_rdoq_ts
int refshim_rdoq_bdpcm( int comp, const int32_t* coef, int w, int h, int bitDepth, int qp, int inputDelta, int intraCu, int dirMode, double lambda, int ctxQp, int ctxInitId,
                        int16_t* q, int32_t* absSum, int32_t* ratesOut )
{
  RefCtx& c = ctx(); (void) c;
  TuRig& r = rig();
  const ComponentID compID = comp ? COMP_Cb : COMP_Y;
  r.setup( w, h, bitDepth, MTS_SKIP, false, intraCu != 0, qp, comp ? CHROMA_444 : CHROMA_400 );
  if( comp ) r.tu.mtsIdx[COMP_Cb] = MTS_SKIP;
  r.cu.bdpcmM[comp ? CH_C : CH_L] = (uint8_t) dirMode;
  r.sps.internalMinusInputBitDepth[CH_L] = inputDelta; r.sps.internalMinusInputBitDepth[CH_C] = inputDelta;
  static thread_local std::unique_ptr<DepQuant> dq;
  if( !dq ) dq.reset( new DepQuant( nullptr, true, false, true ) );
  dq->init( 2, true, 8 );
  dq->m_dLambda = lambda;
  static thread_local std::unique_ptr<Ctx> cabac;
  if( !cabac ) cabac.reset( new Ctx( (const BinProbModel*) nullptr ) );
  cabac->init( ctxQp, ctxInitId );
  QpParam qpp( r.tu, COMP_Y, false );
  CCoeffBuf src( coef, w, w, h );
  TCoeff sum = 0;
  dq->forwardRDPCM( r.tu, compID, src, sum, qpp, *cabac );
  memcpy( q, r.qcoef.data(), sizeof( int16_t ) * w * h );
  *absSum = sum;
  if( ratesOut ) rdoqTsRatesOf( *cabac, ratesOut );
  r.cu.bdpcmM[CH_L] = 0; r.cu.bdpcmM[CH_C] = 0;
  r.sps.internalMinusInputBitDepth[CH_L] = 0; r.sps.internalMinusInputBitDepth[CH_C] = 0;
  return 0;
}

// the same TU through integration/TrQuantB200.h (rateDistOptQuantTSB200); returns 1 when the binding threw
int refshim_rdoq_ts_b200( int comp, const int32_t* coef, int w, int h, int bitDepth, int qp, int inputDelta, int intraCu, double lambda, int ctxQp, int ctxInitId, int16_t* q, int32_t* absSum )
{
  RefCtx& c = ctx(); (void) c;
  TuRig& r = rig();
  const ComponentID compID = comp ? COMP_Cb : COMP_Y;
  r.setup( w, h, bitDepth, MTS_SKIP, false, intraCu != 0, qp, comp ? CHROMA_444 : CHROMA_400 );
  if( comp ) r.tu.mtsIdx[COMP_Cb] = MTS_SKIP;
  r.sps.internalMinusInputBitDepth[CH_L] = inputDelta; r.sps.internalMinusInputBitDepth[CH_C] = inputDelta;
  static thread_local std::unique_ptr<DepQuant> dq;
  if( !dq ) dq.reset( new DepQuant( nullptr, true, false, true ) );
  dq->init( 2, true, 8 );
  dq->m_dLambda = lambda;
  static thread_local std::unique_ptr<Ctx> cabac;
  if( !cabac ) cabac.reset( new Ctx( (const BinProbModel*) nullptr ) );
  cabac->init( ctxQp, ctxInitId );
  QpParam qpp( r.tu, COMP_Y, false );
  CCoeffBuf src( coef, w, w, h );
  TCoeff sum = 0;
  int rc = 0;
  try { rateDistOptQuantTSB200( *dq, tqOfThread(), r.tu, compID, src, sum, qpp, *cabac ); }
  catch( std::exception& e ) { g_b200.error = e.what(); rc = 1; }
  r.sps.internalMinusInputBitDepth[CH_L] = 0; r.sps.internalMinusInputBitDepth[CH_C] = 0;
  if( rc ) return rc;
  memcpy( q, r.qcoef.data(), sizeof( int16_t ) * w * h );
  *absSum = sum;
  return 0;
}

// BDPCM through the bindings: forwardRDPCMB200, and (inverse != 0) Quant::dequant + xITransformSkip of the levels against invTransformNxNB200 (host running sums + the library's
// inverse of skipped transforms); returns 1 when a binding threw
int refshim_rdoq_bdpcm_b200( int comp, const int32_t* coef, int w, int h, int bitDepth, int qp, int inputDelta, int dirMode, double lambda, int ctxQp, int ctxInitId, int16_t* q, int32_t* absSum,
                             int16_t* resiMember, int16_t* resiB200 )
{
  RefCtx& c = ctx(); (void) c;
  TuRig& r = rig();
  const ComponentID compID = comp ? COMP_Cb : COMP_Y;
  r.setup( w, h, bitDepth, MTS_SKIP, false, true, qp, comp ? CHROMA_444 : CHROMA_400 );
  if( comp ) r.tu.mtsIdx[COMP_Cb] = MTS_SKIP;
  r.cu.bdpcmM[comp ? CH_C : CH_L] = (uint8_t) dirMode;
  r.sps.internalMinusInputBitDepth[CH_L] = inputDelta; r.sps.internalMinusInputBitDepth[CH_C] = inputDelta;
  static thread_local std::unique_ptr<DepQuant> dq;
  if( !dq ) dq.reset( new DepQuant( nullptr, true, false, true ) );
  dq->init( 2, true, 8 );
  dq->m_dLambda = lambda;
  static thread_local std::unique_ptr<Ctx> cabac;
  if( !cabac ) cabac.reset( new Ctx( (const BinProbModel*) nullptr ) );
  cabac->init( ctxQp, ctxInitId );
  QpParam qpp( r.tu, COMP_Y, false );
  CCoeffBuf src( coef, w, w, h );
  TCoeff sum = 0;
  int rc = 0;
  try
  {
    forwardRDPCMB200( *dq, tqOfThread(), r.tu, compID, src, sum, qpp, *cabac );
    memcpy( q, r.qcoef.data(), sizeof( int16_t ) * w * h );
    *absSum = sum;
    if( resiMember && resiB200 )
    {
      // inverse path of the levels just produced: the member (Quant::dequant incl. invResDPCM + xITransformSkip) against the binding
      TrQuant& tq = tqOfThread();
      std::vector<Pel> a( (size_t) w * h ), b( (size_t) w * h );
      PelBuf pa( a.data(), w, w, h ), pb( b.data(), w, w, h );
      tq.invTransformNxN( r.tu, compID, pa, qpp );
      invTransformNxNB200( tq, r.tu, compID, pb, qpp );
      memcpy( resiMember, a.data(), sizeof( Pel ) * w * h ); memcpy( resiB200, b.data(), sizeof( Pel ) * w * h );
    }
  }
  catch( std::exception& e ) { g_b200.error = e.what(); rc = 1; }
  r.cu.bdpcmM[CH_L] = 0; r.cu.bdpcmM[CH_C] = 0;
  r.sps.internalMinusInputBitDepth[CH_L] = 0; r.sps.internalMinusInputBitDepth[CH_C] = 0;
  return rc;
}

// scan geometry of DQIntern::Rom for one luma shape, repacked into the 24- / 16-byte records of vvenc_b200/csrc/depquant_core.h (DqScanInfo, DqNbOut);
// fields the reference leaves unset (nextSbbRight / nextSbbBelow off group starts, everything "next" at scan position 0) are reported as 0
int refshim_dep_quant_tables_ex( int chroma, int w, int h, uint8_t* scanInfoOut, uint8_t* nbOutOut );
int refshim_dep_quant_tables( int w, int h, uint8_t* scanInfoOut, uint8_t* nbOutOut ) { return refshim_dep_quant_tables_ex( 0, w, h, scanInfoOut, nbOutOut ); }
int refshim_dep_quant_tables_ex( int chroma, int w, int h, uint8_t* scanInfoOut, uint8_t* nbOutOut )
{
  ctx();
  static thread_local std::unique_ptr<DepQuant> dq;
  if( !dq ) dq.reset( new DepQuant( nullptr, true, false, false ) );
  const CompArea area( chroma ? COMP_Cb : COMP_Y, chroma ? CHROMA_444 : CHROMA_400, Area( 0, 0, w, h ) );
  const DQIntern::TUParameters& tp = *dq->m_scansRom->getTUPars( area, chroma ? COMP_Cb : COMP_Y );
  const int nc = (int) tp.m_numCoeff;
  for( int i = 0; i < nc; i++ )
  {
    const DQIntern::ScanInfo& s = tp.m_scanInfo[i];
    uint8_t* o = scanInfoOut + 24 * i;
    memset( o, 0, 24 );
    const bool first = s.insidePos == 0 && i > 0;
    int16_t v16[4] = { s.rasterPos, s.sbbPos, (int16_t)( first ? s.nextSbbRight : 0 ), (int16_t)( first ? s.nextSbbBelow : 0 ) };
    memcpy( o, v16, 8 );
    int8_t v8[7] = { s.insidePos, (int8_t)( i ? s.nextInsidePos : 0 ), (int8_t) s.spt, s.posX, s.posY, (int8_t)( i ? s.sigCtxOffsetNext : 0 ), (int8_t)( i ? s.gtxCtxOffsetNext : 0 ) };
    memcpy( o + 8, v8, 7 );
    const DQIntern::NbInfoSbb& nb = tp.m_scanId2NbInfoSbb[i];
    o[15] = nb.numInv;
    for( int k = 0; k < 5; k++ ) o[16 + k] = k < nb.numInv ? nb.invInPos[k] : 0;
    const DQIntern::NbInfoOut& no = tp.m_scanId2NbInfoOut[i];
    uint16_t w16[8] = { no.maxDist, no.num, no.outPos[0], no.outPos[1], no.outPos[2], no.outPos[3], no.outPos[4], 0 };
    memcpy( nbOutOut + 16 * i, w16, 16 );
  }
  return nc;
}

// integration/TrQuantB200.h in action: the same TU rig, xT + Quant::quant replaced by xTQuantB200 / invTransformNxN by invTransformNxNB200 on the bound library.
// Return 0, -1 (transform pair not expressible as an mtsIdx) or 1 (the binding threw; text through refshim_b200_error).
int refshim_install_b200_tu( const char* libPath ) { return b200LoadTu( libPath ); }
int refshim_transform_quant_b200( int trHor, int trVer, const int16_t* resi, int stride, int w, int h, int bitDepth, int qp, int isIRAP, int depQuant,
                                  int32_t* coef, int16_t* q, int32_t* absSum, int32_t* lastPos, int32_t* needRdoq )
{
  RefCtx& c = ctx();
  const int mts = mtsIdxFor( trHor, trVer );
  if( mts < 0 ) return -1;
  TuRig& r = rig();
  r.setup( w, h, bitDepth, mts, isIRAP != 0, true, qp );
  r.slice.depQuantEnabled = depQuant != 0;
  CPelBuf  resiBuf( resi, stride, w, h );
  CoeffBuf dst( coef, w, w, h );
  QpParam qpp( r.tu, COMP_Y, false );
  TCoeff sum = 0; bool nr = false;
  try { xTQuantB200( tqOfThread(), r.tu, COMP_Y, resiBuf, dst, qpp, sum, &nr ); }
  catch( std::exception& e ) { g_b200.error = e.what(); r.slice.depQuantEnabled = false; return 1; }
  r.slice.depQuantEnabled = false;
  memcpy( q, r.qcoef.data(), sizeof( int16_t ) * w * h );
  *absSum = sum; *lastPos = r.tu.lastPos[COMP_Y]; *needRdoq = nr ? 1 : 0;
  return 0;
}
int refshim_transform_quant_b200_sdh( int trHor, int trVer, const int16_t* resi, int stride, int w, int h, int bitDepth, int qp, int isIRAP, int depQuant, int signHiding,
                                      int32_t* coef, int16_t* q, int32_t* absSum, int32_t* lastPos, int32_t* needRdoq )
{
  t_rigSignHiding = signHiding != 0;
  const int rc = refshim_transform_quant_b200( trHor, trVer, resi, stride, w, h, bitDepth, qp, isIRAP, depQuant, coef, q, absSum, lastPos, needRdoq );
  t_rigSignHiding = false;
  return rc;
}
int refshim_inv_transform_quant_b200( int trHor, int trVer, const int16_t* q, int w, int h, int bitDepth, int qp, int16_t* resi, int stride )
{
  RefCtx& c = ctx();
  const int mts = mtsIdxFor( trHor, trVer );
  if( mts < 0 ) return -1;
  TuRig& r = rig();
  r.setup( w, h, bitDepth, mts, false, true, qp );
  memcpy( r.qcoef.data(), q, sizeof( int16_t ) * w * h );
  QpParam qpp( r.tu, COMP_Y, false );
  PelBuf out( resi, stride, w, h );
  try { invTransformNxNB200( tqOfThread(), r.tu, COMP_Y, out, qpp ); }
  catch( std::exception& e ) { g_b200.error = e.what(); return 1; }
  return 0;
}

// Inverse path: TrQuant::invTransformNxN (TrQuant.cpp:318-348) = Quant::dequant (Quant.cpp:520) + xIT (TrQuant.cpp:567).
// q: quantised levels [h][w] compact; coef (nullable) receives the dequantised coefficients; resi written with `stride`.
int refshim_inv_transform_quant( int trHor, int trVer, const int16_t* q, int w, int h, int bitDepth, int qp, int32_t* coef, int16_t* resi, int stride )
{
  RefCtx& c = ctx();
  const int mts = mtsIdxFor( trHor, trVer );
  if( mts < 0 ) return -1;
  TuRig& r = rig();
  r.setup( w, h, bitDepth, mts, false, true, qp );
  memcpy( r.qcoef.data(), q, sizeof( int16_t ) * w * h );
  QpParam qpp( r.tu, COMP_Y, false );
  alignas(64) static thread_local TCoeff tmpCoef[ 64 * 64 ];
  CoeffBuf deq( tmpCoef, w, w, h );
  static_cast<Quant*>( tqOfThread().m_quant )->Quant::dequant( r.tu, deq, COMP_Y, qpp );
  if( coef ) memcpy( coef, tmpCoef, sizeof( int32_t ) * w * h );
  PelBuf out( resi, stride, w, h );
  tqOfThread().xIT( r.tu, COMP_Y, CCoeffBuf( tmpCoef, w, w, h ), out );
  return 0;
}

// TrQuant::invTransformNxN of a slice with depQuantEnabled: DepQuant::dequant -> Quantizer::dequantBlock (DepQuant.cpp:1492-1514, 574-629) + xIT.  lastPos: tu.lastPos of the
// TU (the scan position of the last significant level), which dequantBlock starts from.
int refshim_inv_transform_quant_dq( int trHor, int trVer, const int16_t* q, int lastPos, int w, int h, int bitDepth, int qp, int32_t* coef, int16_t* resi, int stride )
{
  RefCtx& c = ctx();
  const int mts = mtsIdxFor( trHor, trVer );
  if( mts < 0 ) return -1;
  TuRig& r = rig();
  r.setup( w, h, bitDepth, mts, false, true, qp );
  r.slice.depQuantEnabled = true;
  r.tu.lastPos[COMP_Y] = lastPos;
  memcpy( r.qcoef.data(), q, sizeof( int16_t ) * w * h );
  QpParam qpp( r.tu, COMP_Y, false );
  alignas(64) static thread_local TCoeff tmpCoef[ 64 * 64 ];
  CoeffBuf deq( tmpCoef, w, w, h );
  tqOfThread().m_quant->dequant( r.tu, deq, COMP_Y, qpp );            // virtual: DepQuant::dequant
  if( coef ) memcpy( coef, tmpCoef, sizeof( int32_t ) * w * h );
  PelBuf out( resi, stride, w, h );
  tqOfThread().xIT( r.tu, COMP_Y, CCoeffBuf( tmpCoef, w, w, h ), out );
  r.slice.depQuantEnabled = false;
  return 0;
}

// the same through integration/TrQuantB200.h (invTransformNxNB200 with par.dep_quant = slice->depQuantEnabled)
int refshim_inv_transform_quant_dq_b200( int trHor, int trVer, const int16_t* q, int lastPos, int w, int h, int bitDepth, int qp, int16_t* resi, int stride )
{
  RefCtx& c = ctx();
  const int mts = mtsIdxFor( trHor, trVer );
  if( mts < 0 ) return -1;
  TuRig& r = rig();
  r.setup( w, h, bitDepth, mts, false, true, qp );
  r.slice.depQuantEnabled = true;
  r.tu.lastPos[COMP_Y] = lastPos;
  memcpy( r.qcoef.data(), q, sizeof( int16_t ) * w * h );
  QpParam qpp( r.tu, COMP_Y, false );
  PelBuf out( resi, stride, w, h );
  int rc = 0;
  try { invTransformNxNB200( tqOfThread(), r.tu, COMP_Y, out, qpp ); }
  catch( std::exception& e ) { g_b200.error = e.what(); rc = 1; }
  r.slice.depQuantEnabled = false;
  return rc;
}

// PelBuf::reconstruct (Buffer.cpp:719) with the slice clipping range of the given bit depth.  The SIMD kernels behind g_pelBufOP use aligned
// loads (the encoder's CU-local buffers are compact and MEMORY_ALIGN_DEF_SIZE aligned), so the probe marshals through such buffers.
struct AlignedPel
{
  int16_t* p;
  explicit AlignedPel( size_t n ) : p( (int16_t*) aligned_alloc( 64, ( n * 2 + 127 ) & ~size_t( 63 ) ) ) {}
  ~AlignedPel() { free( p ); }
};
void refshim_reconstruct( const int16_t* pred, int ps, const int16_t* resi, int rs, int16_t* reco, int cs, int w, int h, int bitDepth )
{
  ClpRng rng; rng.bd = bitDepth;
  AlignedPel a( (size_t) w * h ), b( (size_t) w * h ), d( (size_t) w * h );
  for( int y = 0; y < h; y++ ) { memcpy( a.p + y * w, pred + (ptrdiff_t) y * ps, 2 * w ); memcpy( b.p + y * w, resi + (ptrdiff_t) y * rs, 2 * w ); }
  PelBuf dst( d.p, w, w, h );
  dst.reconstruct( CPelBuf( a.p, w, w, h ), CPelBuf( b.p, w, w, h ), rng );
  for( int y = 0; y < h; y++ ) memcpy( reco + (ptrdiff_t) y * cs, d.p + y * w, 2 * w );
}

// One luma TU candidate end to end the way xIntraCodingTUBlock runs it (IntraSearch.cpp:1353-1429); out4 as in oracle.c orc_tu_roundtrip.
int refshim_tu_roundtrip( int opt, int trHor, int trVer, const int16_t* org, int so, const int16_t* pred, int ps, int w, int h, int bitDepth, int qp, int isIRAP,
                          int16_t* q, int16_t* reco, int cs, uint64_t* out4 )
{
  RefCtx& c = ctx();
  const size_t n = (size_t) w * h;
  AlignedPel ao( n ), ap( n ), resi( n ), rec( n ), zero( n ), arc( n );
  std::vector<int32_t> coef( n + 16 );
  int32_t* coefA = (int32_t*)( ( (uintptr_t) coef.data() + 63 ) & ~uintptr_t( 63 ) );
  for( int y = 0; y < h; y++ ) { memcpy( ao.p + y * w, org + (ptrdiff_t) y * so, 2 * w ); memcpy( ap.p + y * w, pred + (ptrdiff_t) y * ps, 2 * w ); }
  memset( zero.p, 0, 2 * n ); memset( rec.p, 0, 2 * n );
  PelBuf resiBuf( resi.p, w, w, h );
  resiBuf.subtract( CPelBuf( ao.p, w, w, h ), CPelBuf( ap.p, w, w, h ) );
  int32_t absSum = 0, lastPos = 0;
  if( refshim_transform_quant( trHor, trVer, resi.p, w, w, h, bitDepth, qp, isIRAP, coefA, q, &absSum, &lastPos ) ) return -1;
  if( absSum > 0 ) refshim_inv_transform_quant( trHor, trVer, q, w, h, bitDepth, qp, nullptr, rec.p, w );
  refshim_reconstruct( ap.p, w, rec.p, w, arc.p, w, w, h, bitDepth );
  for( int y = 0; y < h; y++ ) memcpy( reco + (ptrdiff_t) y * cs, arc.p + y * w, 2 * w );
  RdCost& rc = c.rd( opt );
  out4[0] = callDist( rc, 0, ao.p, w, arc.p, w, w, h, bitDepth, 0 );
  out4[1] = callDist( rc, 0, resi.p, w, rec.p, w, w, h, bitDepth, 0 );
  out4[2] = callDist( rc, 0, zero.p, w, resi.p, w, w, h, bitDepth, 0 );
  out4[3] = (uint64_t)(uint32_t) absSum | ( (uint64_t)(uint32_t) lastPos << 32 );
  return 0;
}

void refshim_tu_roundtrip_batch( int opt, int trHor, int trVer, const int16_t* org, const int16_t* pred, int n, int w, int h, int bitDepth, int qp, int isIRAP,
                                 int16_t* q, int16_t* reco, uint64_t* out4, int nthreads )
{
  ctx();
  parallelFor( n, nthreads, [&]( int b, int e, int )
  {
    for( int i = b; i < e; i++ )
    {
      const size_t o = (size_t) i * w * h;
      refshim_tu_roundtrip( opt, trHor, trVer, org + o, w, pred + o, w, w, h, bitDepth, qp, isIRAP, q + o, reco + o, w, out4 + 4 * (size_t) i );
    }
  } );
}

// Batch of equal-shape TUs laid out back to back (resi: n * h * w int16, compact), threaded; used by the CPU baseline.
void refshim_transform_quant_batch( int trHor, int trVer, const int16_t* resi, int n, int w, int h, int bitDepth, int qp, int isIRAP,
                                    int16_t* q, int32_t* absSum, int32_t* lastPos, int nthreads )
{
  ctx();
  parallelFor( n, nthreads, [&]( int b, int e, int )
  {
    std::vector<int32_t> coef( (size_t) w * h );
    for( int i = b; i < e; i++ )
      refshim_transform_quant( trHor, trVer, resi + (size_t) i * w * h, w, w, h, bitDepth, qp, isIRAP, coef.data(),
                               q + (size_t) i * w * h, absSum + i, lastPos + i );
  } );
}

// ---------------------------------------------------------------------------------------------------------
// MCTF block-matching errors (MCTF.h:160-166).  besterror = INT_MAX as in vvenc_unit_test.cpp:1552.
int refshim_mctf_err_int( int opt, const int16_t* org, int orgStride, const int16_t* buf, int bufStride, int w, int h, int besterror )
{
  RefCtx& c = ctx();
  return c.mctf[opt?1:0]->m_motionErrorLumaInt8( org, orgStride, buf, bufStride, w, h, besterror );
}

// tap4: 0 -> 6-tap (m_interpolationFilter8 rows, taps 1..6), 1 -> 4-tap; fx, fy in 0..15 (MCTF.cpp:1138-1160)
int refshim_mctf_err_frac( int opt, int tap4, const int16_t* org, int orgStride, const int16_t* buf, int bufStride, int w, int h,
                           int fx, int fy, int bitDepth, int besterror )
{
  RefCtx& c = ctx();
  const int16_t* xf = tap4 ? MCTF::m_interpolationFilter4[fx] : MCTF::m_interpolationFilter8[fx];
  const int16_t* yf = tap4 ? MCTF::m_interpolationFilter4[fy] : MCTF::m_interpolationFilter8[fy];
  return c.mctf[opt?1:0]->m_motionErrorLumaFrac8[tap4?1:0]( org, orgStride, buf, bufStride, w, h, xf, yf, bitDepth, besterror );
}

void refshim_mctf_filters( int16_t* f8 /*16x8*/, int16_t* f4 /*16x4*/ )
{
  memcpy( f8, MCTF::m_interpolationFilter8, sizeof( int16_t ) * 16 * 8 );
  memcpy( f4, MCTF::m_interpolationFilter4, sizeof( int16_t ) * 16 * 4 );
}

// list form for the CPU baseline: desc[i] = { x, y, mvx, mvy (1/16 pel), w, h }
void refshim_mctf_err_list( int opt, int tap4, const int16_t* orgPlane, int orgStride, const int16_t* bufPlane, int bufStride,
                            const int32_t* desc, int n, int bitDepth, int32_t* out, int nthreads )
{
  RefCtx& c = ctx();
  MCTF* m = c.mctf[opt?1:0];
  parallelFor( n, nthreads, [&]( int b, int e, int )
  {
    for( int i = b; i < e; i++ )
    {
      const int32_t* d = desc + 6 * (size_t) i;
      int dx = d[2], dy = d[3];
      const int fx = dx & 15, fy = dy & 15;
      const int16_t* org = orgPlane + (ptrdiff_t) d[1] * orgStride + d[0];
      if( ( fx | fy ) == 0 )
      {
        dx /= 16; dy /= 16;
        const int16_t* buf = bufPlane + (ptrdiff_t)( d[1] + dy ) * bufStride + d[0] + dx;
        out[i] = m->m_motionErrorLumaInt8( org, orgStride, buf, bufStride, d[4], d[5], INT32_MAX );
      }
      else
      {
        dx >>= 4; dy >>= 4;
        const int16_t* buf = bufPlane + (ptrdiff_t)( d[1] + dy ) * bufStride + d[0] + dx;
        const int16_t* xf = tap4 ? MCTF::m_interpolationFilter4[fx] : MCTF::m_interpolationFilter8[fx];
        const int16_t* yf = tap4 ? MCTF::m_interpolationFilter4[fy] : MCTF::m_interpolationFilter8[fy];
        out[i] = m->m_motionErrorLumaFrac8[tap4?1:0]( org, orgStride, buf, bufStride, d[4], d[5], xf, yf, bitDepth, INT32_MAX );
      }
    }
  } );
}

// ---------------------------------------------------------------------------------------------------------
// MCTF apply stage (MCTF.h:168-170 function pointers; callers MCTF.cpp:1437-1483)
void refshim_mctf_apply_frac( int opt, int tap4, const int16_t* org, int orgStride, int16_t* dst, int dstStride, int w, int h, int fx, int fy, int bitDepth )
{
  RefCtx& c = ctx();
  const int16_t* xf = tap4 ? MCTF::m_interpolationFilter4[fx] : MCTF::m_interpolationFilter8[fx];
  const int16_t* yf = tap4 ? MCTF::m_interpolationFilter4[fy] : MCTF::m_interpolationFilter8[fy];
  c.mctf[opt?1:0]->m_applyFrac[CH_L][tap4?1:0]( org, orgStride, dst, dstStride, w, h, xf, yf, bitDepth );
}

void refshim_mctf_planar_correction( int opt, const int16_t* ref, int refStride, int16_t* dst, int dstStride, int w, int h, int bitDepth, unsigned motionError )
{
  RefCtx& c = ctx();
  ClpRng rng; rng.bd = bitDepth;
  c.mctf[opt?1:0]->m_applyPlanarCorrection( ref, refStride, dst, dstStride, w, h, rng, (uint16_t) motionError );
}

// src / dst: whole planes (origin pointers) with the block at (bx, by); corrected: numRefs compact w*h blocks back to back
void refshim_mctf_apply_block( int opt, const int16_t* srcPlane, int srcStride, int16_t* dstPlane, int dstStride, int planeW, int planeH, int bx, int by, int w, int h, int bitDepth,
                               const int16_t* corrected, int numRefs, const int32_t* verror, const double* refStrengths, double weightScaling, double sigmaSq )
{
  RefCtx& c = ctx();
  ClpRng rng; rng.bd = bitDepth;
  const Pel* cp[2 * VVENC_MCTF_RANGE] = { nullptr, };
  for( int i = 0; i < numRefs; i++ ) cp[i] = corrected + (size_t) i * w * h;
  CPelBuf src( srcPlane, srcStride, planeW, planeH );
  PelBuf  dst( dstPlane, dstStride, planeW, planeH );
  c.mctf[opt?1:0]->m_applyBlock( src, dst, CompArea( COMP_Y, CHROMA_400, Area( bx, by, w, h ) ), rng, cp, numRefs, verror, refStrengths, weightScaling, sigmaSq );
}

double refshim_mctf_calc_var( int opt, const int16_t* org, int orgStride, int w, int h )
{
  RefCtx& c = ctx();
  return c.mctf[opt?1:0]->m_calcVar( org, orgStride, w, h );
}

// Replay of the per-block body of MCTF::xFinalizeBlkLine for luma (MCTF.cpp:1437-1483; that member needs a whole encoder configuration, so its 40 lines
// are replayed here on top of the reference's function pointers).  mv4[i] = { x, y, error, rmsme }.
void refshim_mctf_finalize_block( int opt, const int16_t* orgPlane, int orgStride, const int16_t* const* refs, int refStride, int numRefs, const int32_t* mv4,
                                  int planeW, int planeH, int bx, int by, int w, int h, int bitDepth, int tap4, int planarEnabled, const double* refStrengths,
                                  double weightScaling, double sigmaSq, int16_t* dstPlane, int dstStride )
{
  RefCtx& c = ctx();
  MCTF* m = c.mctf[opt?1:0];
  ClpRng rng; rng.bd = bitDepth;
  std::vector<Pel> dstBufs( (size_t) numRefs * w * h + 64 );
  const Pel* cp[2 * VVENC_MCTF_RANGE] = { nullptr, };
  int verror[2 * VVENC_MCTF_RANGE] = { 0, };
  for( int i = 0; i < numRefs; i++ )
  {
    const int32_t* mv = mv4 + 4 * i;
    Pel* dst = dstBufs.data() + (size_t) i * w * h;
    cp[i] = dst;
    const Pel* src = refs[i] + (ptrdiff_t)( by + ( mv[1] >> 4 ) ) * refStride + bx + ( mv[0] >> 4 );
    const int16_t* xf = tap4 ? MCTF::m_interpolationFilter4[mv[0] & 0xf] : MCTF::m_interpolationFilter8[mv[0] & 0xf];
    const int16_t* yf = tap4 ? MCTF::m_interpolationFilter4[mv[1] & 0xf] : MCTF::m_interpolationFilter8[mv[1] & 0xf];
    m->m_applyFrac[CH_L][tap4?1:0]( src, refStride, dst, w, w, h, xf, yf, bitDepth );
    if( mv[3] > 0 && planarEnabled && w == h && w <= 32 )
      m->m_applyPlanarCorrection( orgPlane + (ptrdiff_t) by * orgStride + bx, orgStride, dst, w, w, h, rng, (uint16_t) mv[3] );
    verror[i] = mv[2];
  }
  CPelBuf src( orgPlane, orgStride, planeW, planeH );
  PelBuf  dst( dstPlane, dstStride, planeW, planeH );
  m->m_applyBlock( src, dst, CompArea( COMP_Y, CHROMA_400, Area( bx, by, w, h ) ), rng, cp, numRefs, verror, refStrengths, weightScaling, sigmaSq );
}

// ---------------------------------------------------------------------------------------------------------
// Fractional-pel refinement: the two filter calls behind every filtered block of InterSearch::xPatternRefinement / xExtDIFUpSamplingH/Q
// (InterSearch.cpp:790-850, 2912-3040): filterHor( frac_x, isLast = false ) over h + 7 rows, then filterVer( frac_y, isFirst = false, isLast = true ).
// src points at the integer position of the block; fx, fy in quarter pels.
void refshim_if_two_pass( int opt, const int16_t* src, int srcStride, int w, int h, int fx, int fy, int bitDepth, int reduceTap, int altHpel, int16_t* dst, int dstStride )
{
  // one InterpolationFilter per back end, built once; the fast path takes no lock (this function runs once per candidate on every worker thread of the CPU baseline)
  static std::atomic<InterpolationFilter*> ifs[2];
  InterpolationFilter* fp = ifs[opt?1:0].load( std::memory_order_acquire );
  if( !fp )
  {
    std::lock_guard<std::mutex> lk( g_mtx );
    fp = ifs[opt?1:0].load( std::memory_order_relaxed );
    if( !fp ) { fp = new InterpolationFilter; fp->initInterpolationFilter( opt != 0 ); ifs[opt?1:0].store( fp, std::memory_order_release ); }
  }
  InterpolationFilter& f = *fp;
  ClpRng rng; rng.bd = bitDepth;
  const int ts = w + 8;
  // scratch like the encoder's preallocated m_filteredBlockTmp / m_filteredBlock (InterPrediction.cpp): per thread, grown on demand
  static thread_local std::vector<Pel> tmpStore, outStore;
  if( tmpStore.size() < (size_t)( h + 8 ) * ts + 64 ) tmpStore.resize( (size_t)( h + 8 ) * ts + 64 );
  if( outStore.size() < (size_t) h * ts + 64 ) outStore.resize( (size_t) h * ts + 64 );
  Pel* tmp = (Pel*)( ( (uintptr_t) tmpStore.data() + 63 ) & ~uintptr_t( 63 ) );
  Pel* outp = (Pel*)( ( (uintptr_t) outStore.data() + 63 ) & ~uintptr_t( 63 ) );
  f.filterHor( COMP_Y, src - 3 * srcStride, srcStride, tmp, ts, w, h + 7, fx << 2, false, CHROMA_400, rng, altHpel != 0, 0, reduceTap );
  f.filterVer( COMP_Y, tmp + 3 * ts, ts, outp, ts, w, h, fy << 2, false, true, CHROMA_400, rng, altHpel != 0, 0, reduceTap );
  if( dst ) for( int y = 0; y < h; y++ ) memcpy( dst + (ptrdiff_t) y * dstStride, outp + (ptrdiff_t) y * ts, sizeof( Pel ) * w );
}

// blk[b] = { x, y, w, h, mvx, mvy } ; out[b][j+3][i+3] = distFunc( org, filtered block at quarter-pel offset (i, j) ), family 1 SAD / 2 HAD
void refshim_frac_cost_grid_mt( int opt, const int16_t* orgPlane, int so, const int16_t* refPlane, int sr, const int32_t* blk, int n, int family, int bitDepth, int reduceTap, int altHpel,
                                uint32_t* out, int nthreads );
void refshim_frac_cost_grid( int opt, const int16_t* orgPlane, int so, const int16_t* refPlane, int sr, const int32_t* blk, int n, int family, int bitDepth, int reduceTap, int altHpel, uint32_t* out )
{
  refshim_frac_cost_grid_mt( opt, orgPlane, so, refPlane, sr, blk, n, family, bitDepth, reduceTap, altHpel, out, 1 );
}

// threaded form for the CPU baseline of the row (one RdCost per worker, blocks split statically)
void refshim_frac_cost_grid_mt( int opt, const int16_t* orgPlane, int so, const int16_t* refPlane, int sr, const int32_t* blk, int n, int family, int bitDepth, int reduceTap, int altHpel,
                                uint32_t* out, int nthreads )
{
  ctx();
  refshim_if_two_pass( opt, refPlane, sr, 8, 8, 0, 0, bitDepth, reduceTap, altHpel, nullptr, 0 );   // creates the shared InterpolationFilter before the workers start
  parallelFor( n, nthreads, [&]( int b0, int b1, int )
  {
  RdCost rc; createRd( rc, opt );
  for( int b = b0; b < b1; b++ )
  {
    const int32_t* d = blk + 6 * (size_t) b;
    const int w = d[2], h = d[3];
    AlignedPel org( (size_t) w * h ), pred( (size_t) w * h );
    for( int y = 0; y < h; y++ ) memcpy( org.p + y * w, orgPlane + (ptrdiff_t)( d[1] + y ) * so + d[0], 2 * w );
    for( int j = -3; j <= 3; j++ )
      for( int i = -3; i <= 3; i++ )
      {
        const int16_t* src = refPlane + (ptrdiff_t)( d[1] + d[5] + ( j >> 2 ) ) * sr + d[0] + d[4] + ( i >> 2 );
        refshim_if_two_pass( opt, src, sr, w, h, i & 3, j & 3, bitDepth, reduceTap, altHpel, pred.p, w );
        out[( (size_t) b * 7 + ( j + 3 ) ) * 7 + ( i + 3 )] = (uint32_t) callDist( rc, family, org.p, w, pred.p, w, w, h, bitDepth, 0 );
      }
  }
  } );
}

// ---------------------------------------------------------------------------------------------------------
// The reference's OWN MCTF motion search for one pyramid level: MCTF::motionEstimationLuma (MCTF.cpp:1329-1397) -> estimateLumaLn (:1166-1327),
// run on caller-provided pictures (sample (0,0) pointers into padded buffers).  prev (nullable): motion field of the coarser level
// [prevH][prevW] x { x, y }.  out: [blocksY][blocksX] x { x, y, error, rmsme } ; overlapOut (nullable): doubles.
// Only members the search reads are set (private members are reachable through the probe's `#define private public`).
void refshim_mctf_estimate_level( int opt, const int16_t* org, int orgStride, const int16_t* buf, int bufStride, int width, int height, int bitDepth,
                                  int blockSize, const int32_t* prev, int prevW, int prevH, int factor, int doubleRes, int lowResFilter, int unitSize,
                                  int32_t* out, double* overlapOut )
{
  RefCtx& c = ctx();
  MCTF* m = c.mctf[opt?1:0];
  static VVEncCfg cfg;
  cfg.m_internalBitDepth[CH_L] = bitDepth; cfg.m_internalBitDepth[CH_C] = bitDepth;
  m->m_encCfg = &cfg; m->m_threadPool = nullptr; m->m_searchPttrn = 0; m->m_mctfUnitSize = unitSize; m->m_lowResFltSearch = lowResFilter != 0;
  PelStorage orig, buffer;
  orig.createFromBuf( PelUnitBuf( CHROMA_400, PelBuf( const_cast<Pel*>( org ), orgStride, width, height ) ) );
  buffer.createFromBuf( PelUnitBuf( CHROMA_400, PelBuf( const_cast<Pel*>( buf ), bufStride, width, height ) ) );
  const int bxN = width / blockSize, byN = height / blockSize;                       // MCTF.cpp:688: Array2D<MotionVector>( width / blockSize, height / blockSize )
  Array2D<MotionVector> mvs( bxN, byN );
  Array2D<MotionVector> previous;
  if( prev )
  {
    previous.allocate( prevW, prevH );
    for( int y = 0; y < prevH; y++ ) for( int x = 0; x < prevW; x++ ) { MotionVector& v = previous.get( x, y ); v.x = prev[2 * ( y * prevW + x )]; v.y = prev[2 * ( y * prevW + x ) + 1]; }
  }
  m->motionEstimationLuma( mvs, orig, buffer, blockSize, prev ? &previous : nullptr, factor, doubleRes != 0 );
  for( int y = 0; y < byN; y++ )
    for( int x = 0; x < bxN; x++ )
    {
      const MotionVector& v = mvs.get( x, y );
      int32_t* o = out + 4 * ( y * bxN + x );
      o[0] = v.x; o[1] = v.y; o[2] = v.error; o[3] = v.rmsme;
      if( overlapOut ) overlapOut[y * bxN + x] = v.overlap;
    }
}

// integration/MCTFB200.h in action: the set-up of refshim_mctf_estimate_level with the search pattern exposed; useB200 != 0 runs motionEstimationLumaB200 on the
// bound C-ABI library instead of the member.  Returns 0, or 1 when the binding threw (text through refshim_b200_error).
int refshim_install_b200_mctf( const char* libPath ) { return b200LoadMctf( libPath ); }
int refshim_mctf_estimate_level_b200( int opt, int useB200, const int16_t* org, int orgStride, const int16_t* buf, int bufStride, int width, int height, int bitDepth,
                                      int blockSize, const int32_t* prev, int prevW, int prevH, int factor, int doubleRes, int lowResFilter, int unitSize, int searchPattern,
                                      int32_t* out, double* overlapOut )
{
  RefCtx& c = ctx();
  MCTF* m = c.mctf[opt?1:0];
  static VVEncCfg cfg;
  cfg.m_internalBitDepth[CH_L] = bitDepth; cfg.m_internalBitDepth[CH_C] = bitDepth;
  m->m_encCfg = &cfg; m->m_threadPool = nullptr; m->m_searchPttrn = searchPattern; m->m_mctfUnitSize = unitSize; m->m_lowResFltSearch = lowResFilter != 0;
  PelStorage orig, buffer;
  orig.createFromBuf( PelUnitBuf( CHROMA_400, PelBuf( const_cast<Pel*>( org ), orgStride, width, height ) ) );
  buffer.createFromBuf( PelUnitBuf( CHROMA_400, PelBuf( const_cast<Pel*>( buf ), bufStride, width, height ) ) );
  const int bxN = width / blockSize, byN = height / blockSize;
  Array2D<MotionVector> mvs( bxN + 1, byN + 1 );                                       // room for a partial last block (>= 8 pels)
  Array2D<MotionVector> previous;
  if( prev )
  {
    previous.allocate( prevW, prevH );
    for( int y = 0; y < prevH; y++ ) for( int x = 0; x < prevW; x++ ) { MotionVector& v = previous.get( x, y ); v.x = prev[2 * ( y * prevW + x )]; v.y = prev[2 * ( y * prevW + x ) + 1]; }
  }
  try
  {
    if( useB200 ) motionEstimationLumaB200( *m, mvs, orig, buffer, blockSize, prev ? &previous : nullptr, factor, doubleRes != 0 );
    else          m->motionEstimationLuma( mvs, orig, buffer, blockSize, prev ? &previous : nullptr, factor, doubleRes != 0 );
  }
  catch( std::exception& e ) { g_b200.error = e.what(); m->m_searchPttrn = 0; return 1; }
  m->m_searchPttrn = 0;
  const int oxN = ( width + blockSize - 8 ) / blockSize, oyN = ( height + blockSize - 8 ) / blockSize;     // blocks with x + 8 <= width
  for( int y = 0; y < oyN; y++ )
    for( int x = 0; x < oxN; x++ )
    {
      const MotionVector& v = mvs.get( x, y );
      int32_t* o = out + 4 * ( y * oxN + x );
      o[0] = v.x; o[1] = v.y; o[2] = v.error; o[3] = v.rmsme;
      if( overlapOut ) overlapOut[y * oxN + x] = doubleRes ? v.overlap : 0.0;
    }
  return 0;
}

// The whole MCTF motion search of one neighbour picture, as MCTF::motionEstimationMCTF chains it (MCTF.cpp:666-724): subsampleLuma pyramids (:1072-1097,
// reference code, incl. its border extension) and four or five motionEstimationLuma levels.  org / ref: compact width x height pictures; the probe pads
// them by MCTF_PADDING with border replication as Picture buffers are.  out: [hInBlks][wInBlks] x { x, y, error, rmsme }.
void refshim_mctf_estimate_pyramid( int opt, const int16_t* org, const int16_t* ref, int width, int height, int bitDepth, int unitSize, int addLevel,
                                    int searchPattern, int lowResFilter, int32_t* out )
{
  RefCtx& c = ctx();
  MCTF* m = c.mctf[opt?1:0];
  static VVEncCfg cfg;
  cfg.m_internalBitDepth[CH_L] = bitDepth; cfg.m_internalBitDepth[CH_C] = bitDepth;
  m->m_encCfg = &cfg; m->m_threadPool = nullptr; m->m_searchPttrn = searchPattern; m->m_mctfUnitSize = unitSize; m->m_lowResFltSearch = lowResFilter != 0;
  const int pad = MCTF_PADDING;
  auto load = [&]( PelStorage& ps, const int16_t* src )
  {
    ps.create( CHROMA_400, Area( 0, 0, width, height ), 0, pad );
    PelBuf b = ps.Y();
    for( int y = 0; y < height; y++ ) memcpy( b.buf + (ptrdiff_t) y * b.stride, src + (size_t) y * width, sizeof( Pel ) * width );
    ps.extendBorderPel( pad, pad );
  };
  PelStorage origBuf, refBuf, o2, o4, o8, r2, r4, r8;
  load( origBuf, org ); load( refBuf, ref );
  m->subsampleLuma( origBuf, o2 ); m->subsampleLuma( o2, o4 );
  m->subsampleLuma( refBuf, r2 );  m->subsampleLuma( r2, r4 );
  Array2D<MotionVector> mv_0( width / ( unitSize * 8 ) + 1, height / ( unitSize * 8 ) + 1 );
  Array2D<MotionVector> mv_1( width / ( unitSize * 4 ) + 1, height / ( unitSize * 4 ) + 1 );
  Array2D<MotionVector> mv_2( width / ( unitSize * 2 ) + 1, height / ( unitSize * 2 ) + 1 );
  if( addLevel )
  {
    Array2D<MotionVector> mv_m( width / ( unitSize * 16 ) + 1, height / ( unitSize * 16 ) + 1 );
    m->subsampleLuma( o4, o8 ); m->subsampleLuma( r4, r8 );
    m->motionEstimationLuma( mv_m, o8, r8, 2 * unitSize );
    m->motionEstimationLuma( mv_0, o4, r4, 2 * unitSize, &mv_m, 2 );
  }
  else m->motionEstimationLuma( mv_0, o4, r4, 2 * unitSize );
  m->motionEstimationLuma( mv_1, o2, r2, 2 * unitSize, &mv_0, 2 );
  m->motionEstimationLuma( mv_2, origBuf, refBuf, 2 * unitSize, &mv_1, 2 );
  const int wInBlks = ( width + unitSize - 1 ) / unitSize, hInBlks = ( height + unitSize - 1 ) / unitSize;
  Array2D<MotionVector> mvs( wInBlks, hInBlks );
  m->motionEstimationLuma( mvs, origBuf, refBuf, unitSize, &mv_2, 1, true );
  for( int y = 0; y < hInBlks; y++ )
    for( int x = 0; x < wInBlks; x++ )
    {
      const MotionVector& v = mvs.get( x, y );
      int32_t* o = out + 4 * ( y * wInBlks + x );
      o[0] = v.x; o[1] = v.y; o[2] = v.error; o[3] = v.rmsme;
    }
  origBuf.destroy(); refBuf.destroy(); o2.destroy(); o4.destroy(); r2.destroy(); r4.destroy();
  if( addLevel ) { o8.destroy(); r8.destroy(); }
}

// ---------------------------------------------------------------------------------------------------------
// The reference's OWN full search: InterSearch::xPatternSearch (InterSearch.cpp:2209-2251) called as a member on a default-constructed InterSearch whose
// only live members are the ones the function reads (m_pcRdCost, m_cDistParam, m_lumaClpRng).  Same block list / output layout as refshim_full_search;
// subShiftMode is passed through RdCost::setDistParam's own rule (RdCost.cpp:187-200): 0 -> no sub-sampling, 2 -> every second row when h > 8.
static void patternSearchProbe( int opt, const int16_t* orgPlane, int orgStride, const int16_t* refPlane, int refStride,
                                const int32_t* blk, int n, int bitDepth, int subShiftMode, double lambda, int costScale, int imvShift, int32_t* out, bool b200 )
{
  for( int i = 0; i < n; i++ )
  {
    const int32_t* b = blk + 10 * (size_t) i;
    RdCost rc; createRd( rc, opt );
    BitDepths bd; bd.recon[CH_L] = bitDepth; bd.recon[CH_C] = bitDepth;
    rc.setLambda( lambda, bd );
    rc.selectMotionLambda();
    rc.setCostScale( costScale );
    rc.setPredictor( Mv( b[8], b[9] ) );
    static thread_local InterSearch* isp = new InterSearch;          // large object: heap, one per thread, never initialised beyond the three members read
    InterSearch& is = *isp;
    is.m_pcRdCost = &rc;
    is.m_lumaClpRng.bd = bitDepth;
    const int w = b[2], h = b[3];
    AlignedPel org( (size_t) w * h );                                   // the pattern key is a compact CU-local buffer in the encoder
    for( int y = 0; y < h; y++ ) memcpy( org.p + y * w, orgPlane + (ptrdiff_t)( b[1] + y ) * orgStride + b[0], 2 * w );
    CPelBuf key( org.p, w, w, h );
    InterSearch::TZSearchStruct st;
    memset( &st, 0, sizeof( st ) );
    st.pcPatternKey = &key;
    st.piRefY = refPlane + (ptrdiff_t) b[1] * refStride + b[0];
    st.iRefStride = refStride;
    st.subShiftMode = subShiftMode;
    st.imvShift = (unsigned) imvShift;
    st.searchRange.left = b[4]; st.searchRange.right = b[5]; st.searchRange.top = b[6]; st.searchRange.bottom = b[7];
    Mv mv; Distortion sad = 0;
    if( b200 ) xPatternSearchB200( is, st, mv, sad ); else is.xPatternSearch( st, mv, sad );
    if( sad != st.uiBestSad - rc.getCostOfVectorWithPredictor( mv.hor, mv.ver, st.imvShift ) ) THROW( "ruiSAD is not the best cost minus its MV rate" );
    int32_t* o = out + 4 * (size_t) i;
    o[0] = mv.hor; o[1] = mv.ver; o[2] = (int32_t)( st.uiBestSad & 0xffffffffu ); o[3] = (int32_t)( st.uiBestSad >> 32 );
  }
}
void refshim_pattern_search_member( int opt, const int16_t* orgPlane, int orgStride, const int16_t* refPlane, int refStride,
                                    const int32_t* blk, int n, int bitDepth, int subShiftMode, double lambda, int costScale, int imvShift, int32_t* out )
{
  patternSearchProbe( opt, orgPlane, orgStride, refPlane, refStride, blk, n, bitDepth, subShiftMode, lambda, costScale, imvShift, out, false );
}

// ---------------------------------------------------------------------------------------------------------
// The reference's OWN TZ search: InterSearch::xTZSearch (InterSearch.cpp:2297-2573; diamond / raster / star refinement) called as a member, and the same call through
// xTZSearchB200 (integration/InterSearchB200.h: one dense SAD table, the unmodified member walking it).  The CodingUnit carries what the member reads: position and size,
// cs->pcv (picture and CTU size for xClipMvSearch / xSetSearchRange).  blk[i] = { x, y, w, h, predHor, predVer } with the predictor in internal (1/16 pel) units, used both
// as RdCost predictor and as start vector like xMotionEstimation does (:2040-2043, :2104).  out[i] = { mvx, mvy, ruiSAD lo, hi, uiBestSad lo, hi, table hits, misses }.
// seconds spent inside the search calls of the last probes (member: the whole xTZSearch; per-row binding: the walks only, the tables exist already) -- read and reset
// by refshim_tz_search_seconds(): the host-side cost that remains per PU when the SADs come from the device
static double g_tzSearchSeconds = 0.0;
double refshim_tz_search_seconds() { const double v = g_tzSearchSeconds; g_tzSearchSeconds = 0.0; return v; }
static int tzSearchProbe( int opt, const int16_t* orgPlane, int orgStride, const int16_t* refPlane, int refStride, int picW, int picH, int refReach, const int32_t* blk, int n,
                          int bitDepth, int subShiftMode, double lambda, int searchRange, int ctuSize, int extended, int fast, int integerET, int firstSearchStop, int imvShift,
                          int64_t* out, int b200 /* 0 member, 1 xTZSearchB200 per PU, 2 B200RowSearch: one launch per block size, then the walks */ )
{
  static thread_local InterSearch* isp = new InterSearch;
  static thread_local BlkUniMvInfoBuffer* uni = new BlkUniMvInfoBuffer;
  static thread_local TuRig* rg = new TuRig;
  static thread_local VVEncCfg cfg;
  InterSearch& is = *isp;
  cfg.m_ifpLines = 0; cfg.m_bIntegerET = integerET != 0; cfg.m_bFastMEAssumingSmootherMVEnabled = firstSearchStop != 0;
  is.m_pcEncCfg = &cfg; is.m_iSearchRange = searchRange; is.m_BlkUniMvInfoBuffer = uni; is.m_lumaClpRng.bd = bitDepth;
  TuRig& r = *rg;
  r.setup( 8, 8, bitDepth, 0, false, false, 32 );
  r.sps.CTUSize = ctuSize; r.sps.log2MinCodingBlockSize = 2;
  r.pps.picWidthInLumaSamples = picW; r.pps.picHeightInLumaSamples = picH;
  const unsigned maxQt[3] = { (unsigned) ctuSize, (unsigned) ctuSize, (unsigned) ctuSize };
  PreCalcValues pcv( r.sps, r.pps, maxQt );
  r.cs.pcv = &pcv;
  int rcAll = 0;
  B200RowSearch rows;
  if( b200 == 2 )
  {
    try
    {
      RdCost rc; createRd( rc, opt );
      BitDepths bd; bd.recon[CH_L] = bitDepth; bd.recon[CH_C] = bitDepth;
      rc.setLambda( lambda, bd ); rc.selectMotionLambda(); rc.setCostScale( 2 );
      rows.setPictures( CPelBuf( orgPlane, orgStride, picW, picH ), CPelBuf( refPlane, refStride, picW, picH ), refReach, bitDepth );
      for( int i = 0; i < n; i++ )
      {
        const int32_t* b = blk + 6 * (size_t) i;
        Mv pred( b[4], b[5] ), predQuarter = pred; predQuarter.changePrecision( MV_PRECISION_INTERNAL, MV_PRECISION_QUARTER );
        rows.addTz( b[0], b[1], b[2], b[3], pred, predQuarter, searchRange, fast != 0, refReach );
      }
      rows.runTables( rc, (unsigned) imvShift, subShiftMode );
    }
    catch( std::exception& e ) { g_b200.error = e.what(); r.cs.pcv = nullptr; return 1; }
  }
  for( int i = 0; i < n; i++ )
  {
    const int32_t* b = blk + 6 * (size_t) i;
    const int w = b[2], h = b[3];
    static_cast<UnitArea&>( r.cu ) = UnitArea( CHROMA_400, Area( b[0], b[1], w, h ) );
    RdCost rc; createRd( rc, opt );
    BitDepths bd; bd.recon[CH_L] = bitDepth; bd.recon[CH_C] = bitDepth;
    rc.setLambda( lambda, bd ); rc.selectMotionLambda();
    Mv pred( b[4], b[5] );
    Mv predQuarter = pred; predQuarter.changePrecision( MV_PRECISION_INTERNAL, MV_PRECISION_QUARTER );
    rc.setPredictor( predQuarter ); rc.setCostScale( 2 );
    is.m_pcRdCost = &rc;
    AlignedPel org( (size_t) w * h );
    for( int y = 0; y < h; y++ ) memcpy( org.p + y * w, orgPlane + (ptrdiff_t)( b[1] + y ) * orgStride + b[0], 2 * w );
    CPelBuf key( org.p, w, w, h );
    InterSearch::TZSearchStruct st;
    memset( &st, 0, sizeof( st ) );
    st.pcPatternKey = &key;
    st.piRefY = refPlane + (ptrdiff_t) b[1] * refStride + b[0];
    st.iRefStride = refStride;
    st.subShiftMode = subShiftMode; st.imvShift = (unsigned) imvShift;
    st.uiBestSad = MAX_DISTORTION;
    is.xSetSearchRange( r.cu, pred, searchRange, st.searchRange );                    // :2002 (xMotionEstimation)
    Mv mv = pred; Distortion sad = 0;
    int64_t* o = out + 8 * (size_t) i;
    const auto t0 = std::chrono::steady_clock::now();
    try
    {
      if( b200 == 2 ) { rows.tzSearch( i, is, r.cu, REF_PIC_LIST_0, 0, st, mv, sad, extended != 0, fast != 0 ); o[6] = (int64_t) t_b200tz.hits; o[7] = (int64_t) t_b200tz.misses; }
      else if( b200 ) { xTZSearchB200( is, r.cu, REF_PIC_LIST_0, 0, st, mv, sad, extended != 0, fast != 0, refReach ); o[6] = (int64_t) t_b200tz.hits; o[7] = (int64_t) t_b200tz.misses; }
      else       { is.xTZSearch( r.cu, REF_PIC_LIST_0, 0, st, mv, sad, extended != 0, fast != 0 ); o[6] = o[7] = 0; }
    }
    catch( std::exception& e ) { g_b200.error = e.what(); rcAll = 1; }
    g_tzSearchSeconds += std::chrono::duration<double>( std::chrono::steady_clock::now() - t0 ).count();
    o[0] = mv.hor; o[1] = mv.ver; o[2] = (int64_t) sad; o[3] = 0; o[4] = (int64_t) st.uiBestSad; o[5] = st.uiBestDistance;
  }
  r.cs.pcv = nullptr;
  return rcAll;
}
int refshim_tz_search_member( int opt, const int16_t* orgPlane, int orgStride, const int16_t* refPlane, int refStride, int picW, int picH, int refReach, const int32_t* blk, int n,
                              int bitDepth, int subShiftMode, double lambda, int searchRange, int ctuSize, int extended, int fast, int integerET, int firstSearchStop, int imvShift, int64_t* out )
{
  return tzSearchProbe( opt, orgPlane, orgStride, refPlane, refStride, picW, picH, refReach, blk, n, bitDepth, subShiftMode, lambda, searchRange, ctuSize, extended, fast, integerET,
                        firstSearchStop, imvShift, out, 0 );
}
int refshim_tz_search_b200( int opt, const int16_t* orgPlane, int orgStride, const int16_t* refPlane, int refStride, int picW, int picH, int refReach, const int32_t* blk, int n,
                            int bitDepth, int subShiftMode, double lambda, int searchRange, int ctuSize, int extended, int fast, int integerET, int firstSearchStop, int imvShift, int64_t* out )
{
  return tzSearchProbe( opt, orgPlane, orgStride, refPlane, refStride, picW, picH, refReach, blk, n, bitDepth, subShiftMode, lambda, searchRange, ctuSize, extended, fast, integerET,
                        firstSearchStop, imvShift, out, 1 );
}
// worker-thread form: the PU list is split over nthreads threads, each with its own InterSearch / RdCost / rig and -- through b200CtxOfThread() -- its own vvb_ctx and
// look-up table (thread_local in the binding), as the encoder's pool workers would call it (EncSlice.cpp:142-147)
int refshim_tz_search_b200_mt( int opt, const int16_t* orgPlane, int orgStride, const int16_t* refPlane, int refStride, int picW, int picH, int refReach, const int32_t* blk, int n,
                               int bitDepth, int subShiftMode, double lambda, int searchRange, int ctuSize, int extended, int fast, int integerET, int firstSearchStop, int imvShift,
                               int64_t* out, int nthreads )
{
  std::atomic<int> bad( 0 );
  parallelFor( n, nthreads, [&]( int b, int e, int )
  {
    if( tzSearchProbe( opt, orgPlane, orgStride, refPlane, refStride, picW, picH, refReach, blk + 6 * (size_t) b, e - b, bitDepth, subShiftMode, lambda, searchRange, ctuSize, extended, fast,
                       integerET, firstSearchStop, imvShift, out + 8 * (size_t) b, 1 ) ) bad++;
  } );
  return bad.load() ? 1 : 0;
}
int refshim_tz_search_rows_b200( int opt, const int16_t* orgPlane, int orgStride, const int16_t* refPlane, int refStride, int picW, int picH, int refReach, const int32_t* blk, int n,
                                 int bitDepth, int subShiftMode, double lambda, int searchRange, int ctuSize, int extended, int fast, int integerET, int firstSearchStop, int imvShift, int64_t* out )
{
  return tzSearchProbe( opt, orgPlane, orgStride, refPlane, refStride, picW, picH, refReach, blk, n, bitDepth, subShiftMode, lambda, searchRange, ctuSize, extended, fast, integerET,
                        firstSearchStop, imvShift, out, 2 );
}

// ---------------------------------------------------------------------------------------------------------
// integration/InterSearchB200.h in action: the same set-up as the member probes, the loops replaced by the batched C-ABI calls.  The library is whatever
// refshim_install_b200_search() bound: libvvenc_b200.so on the GPU box, tests/mock (the C ABI answered by the CPU oracle) for the host-logic tests.
int refshim_install_b200_search( const char* libPath ) { return b200LoadSearch( libPath ); }

// 0 = ok; 1 = the binding threw (text through refshim_b200_error)
int refshim_pattern_search_b200( int opt, const int16_t* orgPlane, int orgStride, const int16_t* refPlane, int refStride,
                                 const int32_t* blk, int n, int bitDepth, int subShiftMode, double lambda, int costScale, int imvShift, int32_t* out )
{
  try { patternSearchProbe( opt, orgPlane, orgStride, refPlane, refStride, blk, n, bitDepth, subShiftMode, lambda, costScale, imvShift, out, true ); }
  catch( std::exception& e ) { g_b200.error = e.what(); return 1; }
  return 0;
}

// B200RowSearch: every block of the list queued against the two whole pictures, one launch per block size.  Same block list / output as above.
int refshim_row_search_b200( int opt, const int16_t* orgPlane, int orgStride, const int16_t* refPlane, int refStride, int width, int height, int margin,
                             const int32_t* blk, int n, int bitDepth, int subShiftMode, double lambda, int costScale, int imvShift, int32_t* out )
{
  try
  {
    RdCost rc; createRd( rc, opt );
    BitDepths bd; bd.recon[CH_L] = bitDepth; bd.recon[CH_C] = bitDepth;
    rc.setLambda( lambda, bd ); rc.selectMotionLambda(); rc.setCostScale( costScale );
    B200RowSearch rows;
    rows.setPictures( CPelBuf( orgPlane, orgStride, width, height ), CPelBuf( refPlane, refStride, width, height ), margin, bitDepth );
    for( int i = 0; i < n; i++ )
    {
      const int32_t* b = blk + 10 * (size_t) i;
      InterSearch::SearchRange sr; sr.left = b[4]; sr.right = b[5]; sr.top = b[6]; sr.bottom = b[7];
      rows.add( b[0], b[1], b[2], b[3], sr, Mv( b[8], b[9] ) );
    }
    rows.run( rc, (unsigned) imvShift, subShiftMode );
    for( int i = 0; i < n; i++ )
    {
      const B200RowSearch::Result& r = rows.results()[i];
      int32_t* o = out + 4 * (size_t) i;
      o[0] = r.mv.hor; o[1] = r.mv.ver; o[2] = (int32_t)( r.cost & 0xffffffffu ); o[3] = (int32_t)( r.cost >> 32 );
    }
  }
  catch( std::exception& e ) { g_b200.error = e.what(); return 1; }
  return 0;
}

// The reference's OWN apply stage for a whole luma picture: MCTF::bilateralFilter (MCTF.cpp:1489-1556) -> xFinalizeBlkLine (:1399-1487), called as members.
// org: compact width x height; refs[i]: compact width x height neighbour pictures (padded here by MCTF_PADDING with border replication);
// mv4: [numRefs][hInBlks][wInBlks] x { x, y, error, rmsme }; refIndex[i] = min(5, |POC distance| - 1) selects m_refStrengths[picReordering ? 0 : 1][.].
// strengthsOut (nullable) receives the strengths the reference used, sigmaSqOut its luma sigma^2 -- so that callers of the C ABI can be fed the same numbers.
// org / refs[i] / out: compact planes; with chroma420 the buffers are I420 (Y width x height, then Cb and Cr, (width/2) x (height/2) each) and all components are filtered
static void bilateralFilterProbe( int opt, const int16_t* org, const int16_t* const* refs, int numRefs, const int32_t* mv4, const int32_t* refIndex,
                                  int width, int height, int bitDepth, int unitSize, int qp, double overallStrength, int picReordering, int lowResApply,
                                  int16_t* out, double* strengthsOut, double* sigmaSqOut, bool b200, int chroma420 = 0 )
{
  RefCtx& c = ctx();
  MCTF* m = c.mctf[opt?1:0];
  static VVEncCfg cfg;
  const ChromaFormat chFmt = chroma420 ? CHROMA_420 : CHROMA_400;
  cfg.m_internalBitDepth[CH_L] = bitDepth; cfg.m_internalBitDepth[CH_C] = bitDepth;
  cfg.m_internChromaFormat = chroma420 ? VVENC_CHROMA_420 : VVENC_CHROMA_400; cfg.m_QP = qp; cfg.m_picReordering = picReordering != 0;
  m->m_encCfg = &cfg; m->m_threadPool = nullptr; m->m_mctfUnitSize = unitSize; m->m_lowResFltApply = lowResApply != 0;
  const int pad = MCTF_PADDING;
  const int nComp = chroma420 ? 3 : 1;
  auto load = [&]( PelStorage& ps, const int16_t* src, int margin )
  {
    ps.create( chFmt, Area( 0, 0, width, height ), 0, margin );
    for( int comp = 0; comp < nComp; comp++ )
    {
      PelBuf b = ps.bufs[comp];
      for( int y = 0; y < (int) b.height; y++ ) memcpy( b.buf + (ptrdiff_t) y * b.stride, src + (size_t) y * b.width, sizeof( Pel ) * b.width );
      src += (size_t) b.width * b.height;
    }
    if( margin ) ps.extendBorderPel( margin, true );
  };
  PelStorage orgPic, newOrgPic;
  load( orgPic, org, 0 ); load( newOrgPic, org, 0 );
  const int wInBlks = ( width + unitSize - 1 ) / unitSize, hInBlks = ( height + unitSize - 1 ) / unitSize;
  std::deque<TemporalFilterSourcePicInfo> info( numRefs );
  for( int i = 0; i < numRefs; i++ )
  {
    load( info[i].picBuffer, refs[i], pad );
    info[i].mvs.allocate( wInBlks, hInBlks );
    info[i].index = refIndex[i];
    for( int y = 0; y < hInBlks; y++ )
      for( int x = 0; x < wInBlks; x++ )
      {
        const int32_t* v = mv4 + 4 * ( ( (size_t) i * hInBlks + y ) * wInBlks + x );
        MotionVector& mv = info[i].mvs.get( x, y );
        mv.x = v[0]; mv.y = v[1]; mv.error = v[2]; mv.rmsme = (uint16_t) v[3];
      }
    if( strengthsOut ) strengthsOut[i] = MCTF::m_refStrengths[picReordering ? 0 : 1][refIndex[i]];
  }
  if( sigmaSqOut )
  {
    const double lumaSigmaSq = MCTF::m_sigmaMultiplier * ( 128.0 + 3.0 / 256.0 * qp * qp * qp );
    const double w = 1024.0 / ( ( 1 << bitDepth ) );
    *sigmaSqOut = lumaSigmaSq / ( w * w );
  }
  if( b200 ) bilateralFilterB200( *m, orgPic, info, newOrgPic, overallStrength ); else m->bilateralFilter( orgPic, info, newOrgPic, overallStrength );
  for( int comp = 0; comp < nComp; comp++ )
  {
    CPelBuf r = newOrgPic.bufs[comp];
    for( int y = 0; y < (int) r.height; y++ ) memcpy( out + (size_t) y * r.width, r.buf + (ptrdiff_t) y * r.stride, sizeof( Pel ) * r.width );
    out += (size_t) r.width * r.height;
  }
  orgPic.destroy(); newOrgPic.destroy();
  for( int i = 0; i < numRefs; i++ ) info[i].picBuffer.destroy();
}
void refshim_mctf_bilateral_filter( int opt, const int16_t* org, const int16_t* const* refs, int numRefs, const int32_t* mv4, const int32_t* refIndex,
                                    int width, int height, int bitDepth, int unitSize, int qp, double overallStrength, int picReordering, int lowResApply,
                                    int16_t* out, double* strengthsOut, double* sigmaSqOut )
{
  bilateralFilterProbe( opt, org, refs, numRefs, mv4, refIndex, width, height, bitDepth, unitSize, qp, overallStrength, picReordering, lowResApply, out, strengthsOut, sigmaSqOut, false );
}
// 4:2:0 form: I420 buffers, all three components filtered (chroma: half-size units, vectors scaled by the sub-sampling, MCTF.cpp:1417-1455); useB200 selects the binding
int refshim_mctf_bilateral_filter420( int opt, int useB200, const int16_t* org, const int16_t* const* refs, int numRefs, const int32_t* mv4, const int32_t* refIndex,
                                      int width, int height, int bitDepth, int unitSize, int qp, double overallStrength, int picReordering, int lowResApply, int16_t* out )
{
  try { bilateralFilterProbe( opt, org, refs, numRefs, mv4, refIndex, width, height, bitDepth, unitSize, qp, overallStrength, picReordering, lowResApply, out, nullptr, nullptr, useB200 != 0, 1 ); }
  catch( std::exception& e ) { g_b200.error = e.what(); return 1; }
  return 0;
}
// the same set-up with bilateralFilterB200 (integration/MCTFB200.h) in place of the member; 0 = ok, 1 = the binding threw
int refshim_mctf_bilateral_filter_b200( int opt, const int16_t* org, const int16_t* const* refs, int numRefs, const int32_t* mv4, const int32_t* refIndex,
                                        int width, int height, int bitDepth, int unitSize, int qp, double overallStrength, int picReordering, int lowResApply, int16_t* out )
{
  try { bilateralFilterProbe( opt, org, refs, numRefs, mv4, refIndex, width, height, bitDepth, unitSize, qp, overallStrength, picReordering, lowResApply, out, nullptr, nullptr, true ); }
  catch( std::exception& e ) { g_b200.error = e.what(); return 1; }
  return 0;
}

// The reference's OWN fractional refinement: InterSearch::xPatternSearchFracDIF (InterSearch.cpp:2683-2725) called as a member -- xExtDIFUpSamplingH, the half-pel
// round of xPatternRefinement, xExtDIFUpSamplingQ and the quarter-pel round (fastSubPel 0: all nine positions of both rounds; 1: the preset heuristics, where
// xPatternRefinement filters the half-pel blocks itself and skips positions by s_skipQpelPosition).
// blk[i] = { x, y, w, h, mvx, mvy (integer vector), predHor, predVer } ; out[i] = { halfX, halfY, qterX, qterY, costLo, costHi } with the offsets the member
// returns in rcMvHalf / rcMvQter.  Only the members the call tree reads are initialised (InterPredInterpolation::init allocates the filtered-block buffers).
static void fracSearchProbe( int opt, const int16_t* orgPlane, int orgStride, const int16_t* refPlane, int refStride, const int32_t* blk, int n,
                             int bitDepth, double lambda, int reduceTap, int useHad, int altHpel, int fastSubPel, int32_t* out, bool b200 )
{
  static thread_local InterSearch* isp = nullptr;
  static thread_local int inited = -1;
  static VVEncCfg cfg;
  if( !isp ) isp = new InterSearch;
  InterSearch& is = *isp;
  if( inited != ( opt ? 1 : 0 ) ) { if( inited >= 0 ) is.InterPredInterpolation::destroy(); is.InterPredInterpolation::init( opt != 0 ); inited = opt ? 1 : 0; }
  cfg.m_fastSubPel = fastSubPel; cfg.m_meReduceTap = reduceTap; cfg.m_bUseHADME = useHad != 0; cfg.m_fastHad = useHad == 2;   // useHad: 0 SAD, 1 SATD, 2 fast SATD (DF_HAD_fast, 16x16_fast tiles on square multiples of 32)
  is.m_pcEncCfg = &cfg;
  is.m_lumaClpRng.bd = bitDepth;
  is.m_currChromaFormat = CHROMA_400;
  static CodingUnit dummyCu;
  for( int i = 0; i < n; i++ )
  {
    const int32_t* b = blk + 8 * (size_t) i;
    RdCost rc; createRd( rc, opt );
    BitDepths bd; bd.recon[CH_L] = bitDepth; bd.recon[CH_C] = bitDepth;
    rc.setLambda( lambda, bd );
    rc.selectMotionLambda();
    rc.setPredictor( Mv( b[6], b[7] ) );
    is.m_pcRdCost = &rc;
    const int w = b[2], h = b[3];
    AlignedPel org( (size_t) w * h );
    for( int y = 0; y < h; y++ ) memcpy( org.p + y * w, orgPlane + (ptrdiff_t)( b[1] + y ) * orgStride + b[0], 2 * w );
    CPelBuf key( org.p, w, w, h );
    InterSearch::TZSearchStruct st;
    memset( &st, 0, sizeof( st ) );
    st.pcPatternKey = &key;
    st.piRefY = refPlane + (ptrdiff_t) b[1] * refStride + b[0];
    st.iRefStride = refStride;
    st.imvShift = altHpel ? IMV_HPEL : IMV_OFF;                        // the alternative half-pel filter belongs to AMVR half-pel mode: no quarter-pel round (:2712)
    st.useAltHpelIf = altHpel != 0;
    Mv mvInt( b[4], b[5] ), mvHalf, mvQter;
    Distortion cost = 0;
    if( b200 ) xPatternSearchFracDIFB200( is, st, mvInt, mvHalf, mvQter, cost ); else is.xPatternSearchFracDIF( dummyCu, REF_PIC_LIST_0, 0, st, mvInt, mvHalf, mvQter, cost );
    int32_t* o = out + 6 * (size_t) i;
    o[0] = mvHalf.hor; o[1] = mvHalf.ver; o[2] = mvQter.hor; o[3] = mvQter.ver; o[4] = (int32_t)( cost & 0xffffffffu ); o[5] = (int32_t)( cost >> 32 );
  }
}
void refshim_frac_search_member( int opt, const int16_t* orgPlane, int orgStride, const int16_t* refPlane, int refStride, const int32_t* blk, int n,
                                 int bitDepth, double lambda, int reduceTap, int useHad, int altHpel, int fastSubPel, int32_t* out )
{
  fracSearchProbe( opt, orgPlane, orgStride, refPlane, refStride, blk, n, bitDepth, lambda, reduceTap, useHad, altHpel, fastSubPel, out, false );
}
int refshim_frac_search_b200( int opt, const int16_t* orgPlane, int orgStride, const int16_t* refPlane, int refStride, const int32_t* blk, int n,
                              int bitDepth, double lambda, int reduceTap, int useHad, int altHpel, int32_t* out )
{
  try { fracSearchProbe( opt, orgPlane, orgStride, refPlane, refStride, blk, n, bitDepth, lambda, reduceTap, useHad, altHpel, 0, out, true ); }
  catch( std::exception& e ) { g_b200.error = e.what(); return 1; }
  return 0;
}

// same with the sub-pel control of the presets exposed: fastSubPel 0 (slower) or 1 (fast ... slow); useHad 2 = DF_HAD_fast (m_fastHad)
int refshim_frac_search_b200_ex( int opt, const int16_t* orgPlane, int orgStride, const int16_t* refPlane, int refStride, const int32_t* blk, int n,
                                 int bitDepth, double lambda, int reduceTap, int useHad, int altHpel, int fastSubPel, int32_t* out )
{
  try { fracSearchProbe( opt, orgPlane, orgStride, refPlane, refStride, blk, n, bitDepth, lambda, reduceTap, useHad, altHpel, fastSubPel, out, true ); }
  catch( std::exception& e ) { g_b200.error = e.what(); return 1; }
  return 0;
}

// whole-picture, threaded form of refshim_mctf_finalize_block for the CPU baseline of the apply row: block rows are split over the workers
// (mv4: [numRefs][blocksY * blocksX][4]); planes are sample-(0,0) pointers into padded buffers
void refshim_mctf_finalize_picture( int opt, const int16_t* orgPlane, int orgStride, const int16_t* const* refs, int refStride, int numRefs, const int32_t* mv4,
                                    int width, int height, int blockSize, int bitDepth, int tap4, int planarEnabled, const double* refStrengths,
                                    double weightScaling, double sigmaSq, int16_t* dstPlane, int dstStride, int nthreads )
{
  ctx();
  const int bxN = ( width + blockSize - 1 ) / blockSize, byN = ( height + blockSize - 1 ) / blockSize;
  parallelFor( byN, nthreads, [&]( int r0, int r1, int )
  {
    std::vector<int32_t> mv( (size_t) 4 * numRefs );
    for( int by = r0; by < r1; by++ )
      for( int bx = 0; bx < bxN; bx++ )
      {
        for( int i = 0; i < numRefs; i++ ) memcpy( &mv[4 * i], mv4 + 4 * ( (size_t) i * bxN * byN + (size_t) by * bxN + bx ), 16 );
        const int x = bx * blockSize, y = by * blockSize;
        refshim_mctf_finalize_block( opt, orgPlane, orgStride, refs, refStride, numRefs, mv.data(), width, height, x, y, std::min( blockSize, width - x ), std::min( blockSize, height - y ),
                                     bitDepth, tap4, planarEnabled, refStrengths, weightScaling, sigmaSq, dstPlane, dstStride );
      }
  } );
}

// ---------------------------------------------------------------------------------------------------------
// Affine gradient helpers (AffineGradientSearch.h:67-69)
// opt == 2: an AffineGradientSearch whose pointers installB200() patched (integration/AffineGradientB200.h); needs refshim_install_b200_affine first
int refshim_install_b200_affine( const char* libPath ) { return b200LoadAffine( libPath ); }
static AffineGradientSearch* agsOf( RefCtx& c, int opt )
{
  if( opt != 2 ) return c.ags[opt?1:0];
  if( !c.ags[2] ) { c.ags[2] = new AffineGradientSearch( false ); installB200( *c.ags[2] ); }
  return c.ags[2];
}
void refshim_sobel( int opt, int vertical, const int16_t* pred, int predStride, int16_t* deriv, int derivStride, int w, int h )
{
  RefCtx& c = ctx();
  AffineGradientSearch* a = agsOf( c, opt );
  if( vertical ) a->m_VerticalSobelFilter  ( const_cast<Pel*>( pred ), predStride, deriv, derivStride, w, h );
  else           a->m_HorizontalSobelFilter( const_cast<Pel*>( pred ), predStride, deriv, derivStride, w, h );
}

void refshim_equal_coeff( int opt, int sixParam, const int16_t* resi, int resiStride, const int16_t* dx, const int16_t* dy, int derivStride,
                          int w, int h, int64_t* eq /*7x7, accumulated into*/ )
{
  RefCtx& c = ctx();
  AffineGradientSearch* a = agsOf( c, opt );
  Pel* d[2] = { const_cast<Pel*>( dx ), const_cast<Pel*>( dy ) };
  a->m_EqualCoeffComputer[sixParam?1:0]( const_cast<Pel*>( resi ), resiStride, d, derivStride, w, h, reinterpret_cast<int64_t(*)[7]>( eq ) );
}

} // extern "C"
