// integration/AffineGradientB200.h -- reference-side binding of libvvenc_b200.so for the affine-ME gradient helpers.
//
// AffineGradientSearch keeps three public function pointers (CommonLib/AffineGradientSearch.h:67-69) that its constructor points at the scalar or SIMD
// kernels (AffineGradientSearch.cpp:64-82); installB200( AffineGradientSearch& ) is the counterpart of that selection: Sobel derivative planes and the
// normal-equation accumulation of xAffineMotionEstimation (EncoderLib/InterSearch.cpp:5238) come from vvb_affine_sobel / vvb_affine_equal_coeff.
// One call per invocation (borrowed host blocks), like the RdCost trampolines of RdCostB200.h.  Include after RdCostB200.h and CommonLib/AffineGradientSearch.h.
#pragma once
#include "RdCostB200.h"

struct B200AffineApi
{
  bool bound = false;
  decltype( &vvb_affine_sobel )        sobel = nullptr;
  decltype( &vvb_affine_equal_coeff )  equalCoeff = nullptr;
} ;
static B200AffineApi g_b200a;

inline int b200LoadAffine( const char* libPath )
{
  if( g_b200a.bound ) return 0;
  int rc = b200Load( libPath );
  if( rc ) return rc;
  void* h = g_b200.handle;
#define VVB_RESOLVE( member, name ) g_b200a.member = (decltype( g_b200a.member )) dlsym( h, #name ); if( !g_b200a.member ) { g_b200.error = "missing " #name; return -2; }
  VVB_RESOLVE( sobel, vvb_affine_sobel )  VVB_RESOLVE( equalCoeff, vvb_affine_equal_coeff )
#undef VVB_RESOLVE
  g_b200a.bound = true;
  return 0;
}

template<int VERTICAL> void sobelB200( Pel* const pPred, const int predStride, Pel* const pDerivate, const int derivateBufStride, const int width, const int height )
{
  b200Check( g_b200a.sobel( b200CtxOfThread(), VERTICAL, pPred, predStride, pDerivate, derivateBufStride, width, height ) );
}
template<int SIX_PARAM> void equalCoeffB200( Pel* const pResi, const int resiStride, Pel** const ppDerivate, const int derivateBufStride, const int width, const int height,
                                             int64_t ( *pEqualCoeff )[7] )
{
  b200Check( g_b200a.equalCoeff( b200CtxOfThread(), SIX_PARAM, pResi, resiStride, ppDerivate[0], ppDerivate[1], derivateBufStride, width, height, &pEqualCoeff[0][0] ) );
}

inline void installB200( AffineGradientSearch& ags )
{
  ags.m_HorizontalSobelFilter = sobelB200<0>;
  ags.m_VerticalSobelFilter   = sobelB200<1>;
  ags.m_EqualCoeffComputer[0] = equalCoeffB200<0>;
  ags.m_EqualCoeffComputer[1] = equalCoeffB200<1>;
}
