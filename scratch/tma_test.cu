// Standalone bisect of TMA usage: ./tma_test <mode>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ unsigned smem_u32(const void* p){ return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void init_bar(uint64_t* bar){
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;"::"r"(smem_u32(bar)),"r"(1):"memory");
  asm volatile("fence.proxy.async.shared::cta;":::"memory");
}
__device__ __forceinline__ void wait_bar(uint64_t* bar){
  unsigned done;
  do { asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n":"=r"(done):"r"(smem_u32(bar)),"r"(0):"memory"); } while(!done);
}
__global__ void k2d(const __grid_constant__ CUtensorMap map, float* out, int x0, int x1, int n){
  extern __shared__ __align__(128) float sm[];
  __shared__ uint64_t bar;
  if(threadIdx.x==0) init_bar(&bar);
  __syncthreads();
  if(threadIdx.x==0){
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"::"r"(smem_u32(&bar)),"r"(n*4):"memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(sm)),"l"(reinterpret_cast<uint64_t>(&map)),"r"(smem_u32(&bar)),"r"(x0),"r"(x1):"memory");
  }
  wait_bar(&bar);
  for(int i=threadIdx.x;i<n;i+=blockDim.x) out[i]=sm[i];
}
__global__ void k5d(const __grid_constant__ CUtensorMap map, float* out, int x0, int x1, int x2, int x3, int n){
  extern __shared__ __align__(128) float sm[];
  __shared__ uint64_t bar;
  if(threadIdx.x==0) init_bar(&bar);
  __syncthreads();
  if(threadIdx.x==0){
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"::"r"(smem_u32(&bar)),"r"(n*4):"memory");
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(sm)),"l"(reinterpret_cast<uint64_t>(&map)),"r"(smem_u32(&bar)),"r"(x0),"r"(x1),"r"(x2),"r"(x3),"r"(0):"memory");
  }
  wait_bar(&bar);
  for(int i=threadIdx.x;i<n;i+=blockDim.x) out[i]=sm[i];
}
typedef CUresult (*enc_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc,char**argv){
  int mode=argc>1?atoi(argv[1]):0;
  void* ptr=nullptr; cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled",&ptr,cudaEnableDefault,&q);
  enc_t enc=(enc_t)ptr;
  const int W=64,H=16,D=6,C=32,B=1;
  std::vector<float> h(W*H*D*C*B); for(size_t i=0;i<h.size();i++) h[i]=(float)i;
  float* d; cudaMalloc(&d,h.size()*4); cudaMemcpy(d,h.data(),h.size()*4,cudaMemcpyHostToDevice);
  float* out; cudaMalloc(&out,1<<20);
  CUtensorMap map; cuuint32_t es[5]={1,1,1,1,1};
  cudaError_t e; CUresult r;
  if(mode<4){ // 2D: mode0: box 32x4 at (0,0); mode1: box 12x4 at (0,0); mode2: box 32x4 at (-1,-1); mode3: box 12x4 at(-1,-1)
    cuuint64_t gd[2]={W,(cuuint64_t)H*D*C}; cuuint64_t gs[1]={W*4};
    cuuint32_t bx[2]={(mode&1)?12u:32u,4};
    r=enc(&map,CU_TENSOR_MAP_DATA_TYPE_FLOAT32,2,d,gd,gs,bx,es,CU_TENSOR_MAP_INTERLEAVE_NONE,CU_TENSOR_MAP_SWIZZLE_NONE,CU_TENSOR_MAP_L2_PROMOTION_L2_128B,CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    int n=bx[0]*4; int c=(mode&2)?-1:0;
    int cx=c, cy=c;
    if(argc>3){ cx=atoi(argv[2]); cy=atoi(argv[3]); }
    k2d<<<1,128,n*4>>>(map,out,cx,cy,n);
    e=cudaDeviceSynchronize(); printf("mode %d encode %d run: %s\n",mode,(int)r,cudaGetErrorString(e));
    if(e==cudaSuccess){ std::vector<float> o(n); cudaMemcpy(o.data(),out,n*4,cudaMemcpyDeviceToHost); printf("o[0..3]= %f %f %f %f ; row1: %f %f\n",o[0],o[1],o[2],o[3],o[bx[0]],o[bx[0]+1]); }
  } else { // 5D: mode4: coords 0; mode5: coords -1; mode 6: box inner 32
    cuuint64_t gd[5]={W,H,D,C,B}; cuuint64_t gs[4]={W*4,(cuuint64_t)W*H*4,(cuuint64_t)W*H*D*4,(cuuint64_t)W*H*D*C*4};
    cuuint32_t bx[5]={(mode==6)?32u:12u,4,3,8,1};
    r=enc(&map,CU_TENSOR_MAP_DATA_TYPE_FLOAT32,5,d,gd,gs,bx,es,CU_TENSOR_MAP_INTERLEAVE_NONE,CU_TENSOR_MAP_SWIZZLE_NONE,CU_TENSOR_MAP_L2_PROMOTION_L2_128B,CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    int n=bx[0]*4*3*8; int c=(mode==5)?-1:0;
    k5d<<<1,128,n*4>>>(map,out,c,c,c,4,n);
    e=cudaDeviceSynchronize(); printf("mode %d encode %d run: %s\n",mode,(int)r,cudaGetErrorString(e));
    if(e==cudaSuccess){ std::vector<float> o(n); cudaMemcpy(o.data(),out,n*4,cudaMemcpyDeviceToHost);
      int bad=0; for(int cc=0;cc<8;cc++)for(int z=0;z<3;z++)for(int y=0;y<4;y++)for(int x=0;x<(int)bx[0];x++){ int gx=x+c,gy=y+c,gz=z+c,gc=cc+4; float want=(gx<0||gy<0||gz<0)?0.f:h[((gc*D+gz)*H+gy)*W+gx]; if(o[((cc*3+z)*4+y)*bx[0]+x]!=want) bad++; }
      printf("mismatches %d\n",bad); }
  }
  return 0;
}
