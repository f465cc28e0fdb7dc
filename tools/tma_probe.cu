// stand-alone probe: 2-D TMA tile load of uint16 planes with the descriptor (a) as __grid_constant__ parameter, (b) in global memory
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda/barrier>
namespace cde = cuda::device::experimental;
#include <cstdlib>
__device__ __forceinline__ void mbar_init_s( uint32_t a, uint32_t c ) { asm volatile( "mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"( a ), "r"( c ) : "memory" ); }
__device__ __forceinline__ void mbar_expect( uint32_t a, uint32_t b ) { asm volatile( "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"( a ), "r"( b ) : "memory" ); }
__device__ __forceinline__ void mbar_wait( uint32_t a, uint32_t par )
{
  uint32_t done = 0;
  while( !done ) asm volatile( "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"( done ) : "r"( a ), "r"( par ) : "memory" );
}
template<int MODE>
__global__ void k( const __grid_constant__ CUtensorMap pm, const CUtensorMap* gm, int x, int y, int bw, int bh, uint16_t* out )
{
  extern __shared__ __align__( 128 ) unsigned char sm[];
  // barrier inside the dynamic segment (after the tile) so that no static shared memory shifts the 128-byte aligned tile
  unsigned long long* barp = reinterpret_cast<unsigned long long*>( sm + ( ( bw * bh * 2 + 127 ) & ~127 ) );
  const uint32_t b = (uint32_t) __cvta_generic_to_shared( barp );
  if( threadIdx.x == 0 ) printf( "smem tile addr %u bar %u desc %p\n", (uint32_t) __cvta_generic_to_shared( sm ), b, MODE == 0 ? (const void*) &pm : (const void*) gm );
  if( threadIdx.x == 0 ) { mbar_init_s( b, 1 ); asm volatile( "fence.mbarrier_init.release.cluster;" ::: "memory" ); }
  __syncthreads();
  if( threadIdx.x == 0 )
  {
    const CUtensorMap* d = MODE == 0 ? &pm : gm;
    mbar_expect( b, bw * bh * 2 );
    asm volatile( "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                  :: "r"( (uint32_t) __cvta_generic_to_shared( sm ) ), "l"( (unsigned long long) d ), "r"( b ), "r"( x ), "r"( y ) : "memory" );
  }
  mbar_wait( b, 0 );
  for( int i = threadIdx.x; i < bw * bh; i += blockDim.x ) out[i] = reinterpret_cast<uint16_t*>( sm )[i];
}
// NVIDIA's documented libcu++ form of the same copy (CUDA programming guide, "Using TMA to transfer multi-dimensional arrays")
__global__ void k_lib( const __grid_constant__ CUtensorMap pm, int x, int y, int bw, int bh, uint16_t* out )
{
  extern __shared__ __align__( 128 ) unsigned char sm[];
  __shared__ cuda::barrier<cuda::thread_scope_block> bar;
  if( threadIdx.x == 0 ) { init( &bar, blockDim.x ); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  cuda::barrier<cuda::thread_scope_block>::arrival_token token;
  if( threadIdx.x == 0 )
  {
    cde::cp_async_bulk_tensor_2d_global_to_shared( sm, &pm, x, y, bar );
    token = cuda::device::barrier_arrive_tx( bar, 1, bw * bh * 2 );
  }
  else token = bar.arrive();
  bar.wait( std::move( token ) );
  for( int i = threadIdx.x; i < bw * bh; i += blockDim.x ) out[i] = reinterpret_cast<uint16_t*>( sm )[i];
}

int main( int argc, char** argv )
{
  const int onlyMode = argc > 1 ? atoi( argv[1] ) : -1;
  const int variant = argc > 2 ? atoi( argv[2] ) : 0;   // 0: uint16 + L2_128B, 1: uint16 + promotion none, 2: int32 elements + promotion none
  const int S = 288, R = 224, bw = 56, bh = 36;
  std::vector<uint16_t> h( S * R ); for( int i = 0; i < S * R; i++ ) h[i] = (uint16_t)( i * 7 + 3 );
  uint16_t *d, *o; cudaMalloc( &d, S * R * 2 ); cudaMalloc( &o, bw * bh * 2 ); cudaMemcpy( d, h.data(), S * R * 2, cudaMemcpyHostToDevice );
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint( "cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q );
  printf( "entry point: %s q=%d fn=%p\n", cudaGetErrorString( e ), (int) q, fn );
  typedef CUresult ( *EncodeFn )( CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill );
  CUtensorMap tm; memset( &tm, 0, sizeof( tm ) );
  const cuuint64_t gdim[2] = { (cuuint64_t)( variant == 2 ? S / 2 : S ), R }; const cuuint64_t gstr[1] = { S * 2 };
  const cuuint32_t box[2] = { (cuuint32_t)( variant == 2 ? bw / 2 : bw ), bh }; const cuuint32_t es[2] = { 1, 1 };
  CUresult r = ( (EncodeFn) fn )( &tm, variant == 2 ? CU_TENSOR_MAP_DATA_TYPE_INT32 : variant == 3 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : variant == 4 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, variant == 0 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE );
  printf( "encode: %d\n", (int) r );
  CUtensorMap* gtm; cudaMalloc( &gtm, sizeof( tm ) ); cudaMemcpy( gtm, &tm, sizeof( tm ), cudaMemcpyHostToDevice );
  for( int mode = 0; mode < 3; mode++ )
  {
    if( onlyMode >= 0 && mode != onlyMode ) continue;
    const int x = argc > 3 ? atoi( argv[3] ) : ( variant == 2 ? 16 : 33 ), y = 17;     // variant 2: x counts int32 elements (= pel 32)
    if( mode == 2 ) k_lib<<<1, 128, bw * bh * 2 + 256>>>( tm, x, y, bw, bh, o ); else
    if( mode == 0 ) k<0><<<1, 128, bw * bh * 2 + 256>>>( tm, gtm, x, y, bw, bh, o ); else k<1><<<1, 128, bw * bh * 2 + 256>>>( tm, gtm, x, y, bw, bh, o );
    e = cudaDeviceSynchronize();
    printf( "mode %d: %s\n", mode, cudaGetErrorString( e ) );
    if( e != cudaSuccess ) return 1;
    std::vector<uint16_t> res( bw * bh ); cudaMemcpy( res.data(), o, bw * bh * 2, cudaMemcpyDeviceToHost );
    int bad = 0; for( int r2 = 0; r2 < bh; r2++ ) for( int c = 0; c < bw; c++ ) bad += res[r2 * bw + c] != h[( y + r2 ) * S + ( variant == 2 ? 2 * x : x ) + c];
    printf( "mode %d mismatches %d\n", mode, bad );
  }
  return 0;
}
