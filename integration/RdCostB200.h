// integration/RdCostB200.h -- reference-side binding of libvvenc_b200.so for VVenC's RdCost function-pointer tables.
//
// This is the file a VVenC maintainer would add next to CommonLib/x86/RdCostX86.h: FpDistFunc / FpDistFuncX5-shaped
// trampolines (CommonLib/RdCost.h:74-75) into the C ABI of include/vvenc_b200.h, and installB200(), the counterpart of RdCost::_initRdCostX86()
// (CommonLib/x86/RdCostX86.h:3376-3425) that overwrites row [0] of m_afpDistortFunc, m_afpDistortFuncX5 and m_fxdWtdPredPtr (RdCost.h:117-121).
// The library is bound at run time (dlopen), so the encoder carries no link-time dependency on CUDA.  Errors of the C ABI are turned back into the
// reference's THROW semantics.  One vvb_ctx per calling thread (the encoder owns one RdCost per worker thread, EncoderLib/EncSlice.cpp:142-147).
//
// It is compiled for real against the unmodified reference by oracle/Makefile.ref (oracle/ref_shim.cpp includes it) and exercised on the GPU box by
// tests/test_gpu_dropin.py.  It must be included after CommonLib/RdCost.h inside namespace scope where `vvenc` names are visible.
#pragma once
#include <dlfcn.h>
#include <string>
#include "../include/vvenc_b200.h"

struct B200Api
{
  void* handle = nullptr;
  decltype( &vvb_create )          create = nullptr;
  decltype( &vvb_last_error )      lastError = nullptr;
  decltype( &vvb_dist_block )      distBlock = nullptr;
  decltype( &vvb_sad_mask_block )  sadMask = nullptr;
  decltype( &vvb_sad_x5_block )    sadX5 = nullptr;
  decltype( &vvb_fix_wsse_block )  fixWsse = nullptr;
  decltype( &vvb_launch_count )    launchCount = nullptr;
  std::string error;
} ;
static B200Api g_b200;

// binds the entry points; returns 0, -1 (library not loadable) or -2 (symbol missing), text in g_b200.error
inline int b200Load( const char* libPath )
{
  if( g_b200.handle ) return 0;
  void* h = dlopen( libPath, RTLD_NOW | RTLD_LOCAL );
  if( !h ) { g_b200.error = dlerror(); return -1; }
#define VVB_RESOLVE( member, name ) g_b200.member = (decltype( g_b200.member )) dlsym( h, #name ); if( !g_b200.member ) { g_b200.error = "missing " #name; dlclose( h ); return -2; }
  VVB_RESOLVE( create, vvb_create )  VVB_RESOLVE( lastError, vvb_last_error )  VVB_RESOLVE( distBlock, vvb_dist_block )  VVB_RESOLVE( sadMask, vvb_sad_mask_block )
  VVB_RESOLVE( sadX5, vvb_sad_x5_block )  VVB_RESOLVE( fixWsse, vvb_fix_wsse_block )  VVB_RESOLVE( launchCount, vvb_launch_count )
#undef VVB_RESOLVE
  g_b200.handle = h;
  return 0;
}

static thread_local vvb_ctx* t_b200ctx = nullptr;
inline vvb_ctx* b200CtxOfThread()
{
  if( !t_b200ctx && ( !g_b200.create || g_b200.create( &t_b200ctx, 0 ) != VVB_OK ) ) THROW( "no B200 context" );
  return t_b200ctx;
}

// C-ABI status -> the reference's THROW (used by the batched bindings: InterSearchB200.h, MCTFB200.h, TrQuantB200.h)
inline void b200Check( int rc ) { if( rc != VVB_OK ) THROW( g_b200.lastError( b200CtxOfThread() ) ); }

template<int FAM> Distortion distB200( const DistParam& dp )
{
  if( dp.applyWeight ) THROW( " no support" );
  int err = 0;
  const Distortion d = g_b200.distBlock( b200CtxOfThread(), FAM, dp.org.buf, dp.org.stride, dp.cur.buf, dp.cur.stride,
                                         dp.org.width, dp.org.height, dp.bitDepth, dp.subShift, &err );
  if( err ) THROW( g_b200.lastError( b200CtxOfThread() ) );
  return d;
}
inline Distortion sadMaskB200( const DistParam& dp )
{
  int err = 0;
  const Distortion d = g_b200.sadMask( b200CtxOfThread(), dp.org.buf, dp.org.stride, dp.cur.buf, dp.cur.stride, dp.org.width, dp.org.height,
                                       dp.mask, dp.maskStride, dp.stepX, dp.maskStride2, dp.subShift, &err );
  if( err ) THROW( g_b200.lastError( b200CtxOfThread() ) );
  return d;
}
inline void sadX5B200( const DistParam& dp, Distortion* cost, bool centre )
{
  uint64_t c5[5];
  if( g_b200.sadX5( b200CtxOfThread(), dp.org.buf, dp.org.stride, dp.cur.buf, dp.cur.stride, dp.org.width, dp.org.height, dp.subShift, centre, c5 ) ) THROW( "b200 sadX5" );
  for( int i = 0; i < 5; i++ ) if( i != 2 || centre ) cost[i] = c5[i];
}
inline Distortion fixWsseB200( const DistParam& dp, uint32_t w )
{
  int err = 0;
  const Distortion d = g_b200.fixWsse( b200CtxOfThread(), dp.org.buf, dp.org.stride, dp.cur.buf, dp.cur.stride, dp.org.width, dp.org.height, w, &err );
  if( err ) THROW( g_b200.lastError( b200CtxOfThread() ) );
  return d;
}

inline void installB200( RdCost& rc )      // slot = base + log2(width), TypeDef.h:339-382; row [1] (>10 bit) stays scalar like RdCost.cpp:125-126
{
  for( int l = 1; l < 8; l++ )
  {
    rc.m_afpDistortFunc[0][DF_SSE      + l] = distB200<VVB_DF_SSE>;
    rc.m_afpDistortFunc[0][DF_SAD      + l] = distB200<VVB_DF_SAD>;
    rc.m_afpDistortFunc[0][DF_HAD      + l] = distB200<VVB_DF_HAD>;
    rc.m_afpDistortFunc[0][DF_HAD_fast + l] = distB200<VVB_DF_HAD_FAST>;
  }
  rc.m_afpDistortFunc[0][DF_HAD_2SAD]      = distB200<VVB_DF_HAD_2SAD>;
  rc.m_afpDistortFunc[0][DF_SAD_WITH_MASK] = sadMaskB200;
  rc.m_afpDistortFuncX5[0] = sadX5B200;  rc.m_afpDistortFuncX5[1] = sadX5B200;
  rc.m_fxdWtdPredPtr       = fixWsseB200;
}

