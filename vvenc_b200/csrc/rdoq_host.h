// rdoq_host.h -- host-side set-up of the RDOQ kernel (plain C++, no CUDA): the per-call constants QuantRDOQ2::xRateDistOptQuantFast derives before its loops
// (CommonLib/QuantRDOQ2.cpp:500-559, 573-583) and the error scale of xSetErrScaleCoeffNoScalingList (:203-219, double arithmetic in the reference's operation
// order; build with -ffp-contract=off).
#pragma once
#include "rdoq_core.h"
#include <cmath>

namespace vvbrq {

inline int rq_shape_ok( int w, int h )
{
  int lw = 0, lh = 0;
  while( ( 1 << lw ) < w ) lw++;
  while( ( 1 << lh ) < h ) lh++;
  return ( 1 << lw ) == w && ( 1 << lh ) == h && lw >= 2 && lw <= 6 && lh >= 2 && lh <= 6;
}

// qpInternal = cQP.Qp( false ): clip( CU QP + qpBdOffset ) (QpParam, Quant.cpp:89-124).  lfnst = tu.cu->lfnstIdx > 0 (the routine looks at the CU's index for every
// component, :552-559).  sbtZeroOut = the condition of TransformUnit::getTbAreaAfterCoefZeroOut (Unit.cpp:580): sps.MTS && cu.sbtInfo && w <= 32 && h <= 32 && luma.
inline RqPar rq_init_par( int w, int h, int bitDepth, int qpInternal, int lfnst, int sbtZeroOut, int signHiding, int isChroma, double lambda, int thrVal )
{
  static const int quantScales[2][6] = { { 26214, 23302, 20560, 18396, 16384, 14564 }, { 18396, 16384, 14564, 13107, 11651, 10280 } };      // g_quantScales, Rom.cpp:1390-1394
  int lw = 0, lh = 0;
  while( ( 1 << lw ) < w ) lw++;
  while( ( 1 << lh ) < h ) lh++;
  const int per = qpInternal / 6, rem = qpInternal % 6;
  const int maxLog2TrDynamicRange = 15;
  const int transformShift = maxLog2TrDynamicRange - bitDepth - ( ( lw + lh ) >> 1 );          // getTransformShift, Quant.h:69-72
  const bool sqrt2 = ( ( lw + lh ) & 1 ) != 0;                                                  // TU::needsSqrt2Scale (not transform skipped)
  RqPar p;
  p.width = w; p.height = h; p.log2W = lw;
  p.regionW = w < 32 ? w : 32;
  const int regionH = h < 32 ? h : 32;
  p.numCG = lfnst ? 1 : ( p.regionW * regionH ) >> 4;                                           // :553
  p.firstScanPos = ( p.numCG << 4 ) - 1;                                                        // :554
  if( lfnst && ( ( w == 4 && h == 4 ) || ( w == 8 && h == 8 ) ) ) p.firstScanPos = 7;           // :556-559
  p.quantScale = quantScales[sqrt2 ? 1 : 0][rem];                                               // :518
  p.qBits = 14 + per + transformShift + ( sqrt2 ? -1 : 0 );                                     // :522 (QUANT_SHIFT = 14)
  {
    // xSetErrScaleCoeffNoScalingList, :203-219 (SCALE_BITS = 15, DISTORTION_PRECISION_ADJUSTMENT = 0)
    const double dTransShift = (double) transformShift + ( sqrt2 ? -0.5 : 0.0 );
    double dErrScale = pow( 2.0, ( (double) RQ_SCALE_BITS / 2.0 ) );
    dErrScale = dErrScale * pow( 2.0, ( -( dTransShift ) ) );
    const int QStep = p.quantScale;
    const double errScale = dErrScale / QStep / ( 1 << 0 );
    p.errScale = (int)( errScale * (double)( 1 << RQ_ERR_SCALE_SHIFT ) );
  }
  int32_t thres;                                                                                // :573-583
  if( p.qBits ) thres = (int32_t)( (int64_t) thrVal << ( p.qBits - 1 ) );
  else          thres = (int32_t)( (int64_t)( thrVal >> 1 ) << p.qBits );
  p.useThres = thres / ( p.quantScale << 2 );
  int zw = p.regionW, zh = regionH;                                                             // getTbAreaAfterCoefZeroOut, Unit.cpp:574-589
  if( sbtZeroOut && !isChroma && w <= 32 && h <= 32 ) { if( w == 32 ) zw = 16; if( h == 32 ) zh = 16; }
  p.remRegBins = ( zw * zh * 28 ) >> 4;                                                         // MAX_TU_LEVEL_CTX_CODED_BIN_CONSTRAINT = 28, :538-539
  p.signHiding = signHiding ? 1 : 0;
  p.isChroma = isChroma ? 1 : 0;
  p.pad = 0;
  p.lambda = lambda;
  return p;
}

// second engine: lambda * bits tables (the products rq_icost / rq_level_rate_cost would form per coefficient) and the group raster position -> group scan index map
inline RqCost rq_init_cost( const RqPar& p, const RqRates& r )
{
  RqCost c;
  for( int i = 0; i < 12; i++ ) for( int b = 0; b < 2; b++ ) c.sig[i][b] = rq_icost( p, r.sigBits[i][b] );
  for( int i = 0; i < 21; i++ ) for( int l = 1; l <= 3; l++ ) c.lvl[i][l - 1] = rq_level_rate_cost( p, (uint32_t) l, r.parBits[i], r.gt1Bits[i], r.gt2Bits[i], 4, 0, 0 );
  return c;
}
inline void rq_build_cg_index( const int32_t* scan, int w, int h, uint8_t out[64] )
{
  const int rw = w < 32 ? w : 32, rh = h < 32 ? h : 32, wg = rw >> 2, n = wg * ( rh >> 2 );
  for( int i = 0; i < 64; i++ ) out[i] = 0;
  for( int g = 0; g < n; g++ ) { const int r = scan[g << 4], x = r % rw, y = r / rw; out[( y >> 2 ) * wg + ( x >> 2 )] = (uint8_t) g; }
}

// transform-skip variant (QuantRDOQ::rateDistOptQuantTS, QuantRDOQ.cpp:1156-1161, 1183): qpTs = cQP.Qp( true ) = max( clip( CU QP + qpBdOffset ), 4 + 6 * internalMinusInputBitDepth )
inline int rq_ts_shape_ok( int w, int h ) { return rq_shape_ok( w, h ) && w <= 32 && h <= 32; }
inline RqTsPar rq_ts_init_par( int w, int h, int bitDepth, int qpTs, double lambda )
{
  static const int quantScales[6] = { 26214, 23302, 20560, 18396, 16384, 14564 };                // g_quantScales[0]: no sqrt(2) compensation for skipped transforms
  int lw = 0;
  while( ( 1 << lw ) < w ) lw++;
  RqTsPar p;
  p.width = w; p.height = h; p.log2W = lw;
  p.quantScale = quantScales[qpTs % 6];
  p.qBits = 14 + qpTs / 6;
  p.maxCtxBins = ( w * h * 7 ) >> 2;
  p.pad[0] = p.pad[1] = 0;
  {
    // xGetErrScaleCoeff( false, w, h, rem, 15, bitDepth, true ), QuantRDOQ.cpp:319-329: the transform shift is 0 for skipped transforms
    double dErrScale = (double)( 1 << RQ_SCALE_BITS );
    const double dTransShift = (double) 0 + 0.0;
    dErrScale = dErrScale * pow( 2.0, ( -2.0 * dTransShift ) );
    const int QStep = p.quantScale;
    p.errorScale = dErrScale / QStep / QStep / ( 1 << 0 );
  }
  (void) bitDepth;
  p.lambda = lambda;
  return p;
}

// BDPCM (QuantRDOQ::forwardRDPCM, QuantRDOQ.cpp:1381-1383): the dequantiser of the reconstruction the next position predicts from
inline RqBdpcmPar rq_bdpcm_init_par( int dirMode, int qpTs )
{
  static const int invQuantScales[6] = { 40, 45, 51, 57, 64, 72 };                               // g_invQuantScales[0], Rom.cpp:1396-1400
  RqBdpcmPar b;
  b.dirMode = dirMode; b.dqScale = invQuantScales[qpTs % 6]; b.dqRightShift = 6 - qpTs / 6; b.pad = 0;      // IQUANT_SHIFT = 6
  return b;
}

// scan position -> raster index inside the scanned region (row pitch min( 32, w )): grouped 4x4 up-right diagonal scan (Rom.cpp:1098-1136, 1236-1284)
inline void rq_build_scan( int w, int h, int32_t* out /* min(32,w) * min(32,h) */ )
{
  auto diag = []( int bw, int bh, int* xs, int* ys )
  {
    int line = 0, col = 0;
    for( int i = 0; i < bw * bh; i++ )
    {
      xs[i] = col; ys[i] = line;
      if( col == bw - 1 || line == 0 ) { line += col + 1; col = 0; if( line >= bh ) { col += line - ( bh - 1 ); line = bh - 1; } }
      else { col++; line--; }
    }
  };
  const int rw = w < 32 ? w : 32, rh = h < 32 ? h : 32;
  int cx[16], cy[16], gx[64], gy[64];
  diag( 4, 4, cx, cy );
  diag( rw >> 2, rh >> 2, gx, gy );
  for( int g = 0; g < ( rw >> 2 ) * ( rh >> 2 ); g++ )
    for( int c = 0; c < 16; c++ ) out[g * 16 + c] = ( gy[g] * 4 + cy[c] ) * rw + gx[g] * 4 + cx[c];
}

} // namespace vvbrq
