// Microbenchmark: scalar FFMA vs packed FFMA2 issue rate on sm_100a (one-off design probe).
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ void ffma2(float2& d, float2 a, float2 b){
  unsigned long long dd=*reinterpret_cast<unsigned long long*>(&d), aa=*reinterpret_cast<unsigned long long*>(&a), bb=*reinterpret_cast<unsigned long long*>(&b);
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
  d=*reinterpret_cast<float2*>(&dd);
}
template<int NACC>
__global__ void __launch_bounds__(256) k_ffma(float* out, int iters, float a, float b){
  float acc[NACC];
  #pragma unroll
  for(int j=0;j<NACC;j++) acc[j]=threadIdx.x*0.001f+j;
  float x=a+threadIdx.x*1e-6f, y=b;
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int j=0;j<NACC;j++) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[j]) : "f"(x), "f"(y));
  }
  float s=0; 
  #pragma unroll
  for(int j=0;j<NACC;j++) s+=acc[j];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int NACC>
__global__ void __launch_bounds__(256) k_ffma2(float* out, int iters, float a, float b){
  float2 acc[NACC];
  #pragma unroll
  for(int j=0;j<NACC;j++) acc[j]=make_float2(threadIdx.x*0.001f+j, j);
  float2 x=make_float2(a+threadIdx.x*1e-6f,a+threadIdx.x*1e-6f), y=make_float2(b,b*1.0001f);
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int j=0;j<NACC;j++) ffma2(acc[j],x,y);
  }
  float s=0;
  #pragma unroll
  for(int j=0;j<NACC;j++) s+=acc[j].x+acc[j].y;
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
// conv-like inner loop: weights from smem (broadcast LDS.128), x from smem (LDS.128 per lane), 4 vox x 8 co
template<bool PACKED>
__global__ void __launch_bounds__(256) k_convlike(float* out, int iters){
  __shared__ __align__(16) float sw[27*8*8];
  __shared__ __align__(16) float sx[8*(256*4+8)];
  for(int i=threadIdx.x;i<27*8*8;i+=256) sw[i]=i*1e-4f;
  for(int i=threadIdx.x;i<8*(256*4+8);i+=256) sx[i]=i*1e-5f;
  __syncthreads();
  float2 acc[4][4];
  #pragma unroll
  for(int v=0;v<4;v++) for(int j=0;j<4;j++) acc[v][j]=make_float2(0.f,0.f);
  for(int it=0;it<iters;it++){
    #pragma unroll 1
    for(int c=0;c<8;c++){
      #pragma unroll
      for(int t9=0;t9<9;t9++){
        const float* xp = sx + c*(256*4+8) + threadIdx.x*4 + (t9&1)*4;
        float4 xa=*reinterpret_cast<const float4*>(xp);
        float2 xb=*reinterpret_cast<const float2*>(xp+4);
        float xs[6]={xa.x,xa.y,xa.z,xa.w,xb.x,xb.y};
        #pragma unroll
        for(int kw=0;kw<3;kw++){
          const float4* wp=reinterpret_cast<const float4*>(sw + ((t9*3+kw)*8+c)*8);
          float4 w0=wp[0], w1=wp[1];
          #pragma unroll
          for(int v=0;v<4;v++){
            float xv=xs[v+kw];
            if(PACKED){
              float2 xx=make_float2(xv,xv);
              ffma2(acc[v][0],xx,make_float2(w0.x,w0.y)); ffma2(acc[v][1],xx,make_float2(w0.z,w0.w));
              ffma2(acc[v][2],xx,make_float2(w1.x,w1.y)); ffma2(acc[v][3],xx,make_float2(w1.z,w1.w));
            } else {
              acc[v][0].x=fmaf(xv,w0.x,acc[v][0].x); acc[v][0].y=fmaf(xv,w0.y,acc[v][0].y);
              acc[v][1].x=fmaf(xv,w0.z,acc[v][1].x); acc[v][1].y=fmaf(xv,w0.w,acc[v][1].y);
              acc[v][2].x=fmaf(xv,w1.x,acc[v][2].x); acc[v][2].y=fmaf(xv,w1.y,acc[v][2].y);
              acc[v][3].x=fmaf(xv,w1.z,acc[v][3].x); acc[v][3].y=fmaf(xv,w1.w,acc[v][3].y);
            }
          }
        }
      }
    }
  }
  float s=0;
  #pragma unroll
  for(int v=0;v<4;v++) for(int j=0;j<4;j++) s+=acc[v][j].x+acc[v][j].y;
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<class F> float timeit(F f){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1); return ms;
}
int main(){
  cudaDeviceProp p; cudaGetDeviceProperties(&p,0);
  int sms=p.multiProcessorCount; printf("dev %s sms %d clock %d kHz\n",p.name,sms,p.clockRate);
  float* out; cudaMalloc(&out, sizeof(float)*sms*8*256);
  int iters=20000;
  for(int bps=1;bps<=4;bps*=2){
    float ms=timeit([&]{k_ffma<32><<<sms*bps,256>>>(out,iters,1.0001f,0.9999f);});
    double fma=(double)sms*bps*256*32*iters; printf("FFMA  x32acc bps=%d: %.3f ms  %.2f TFLOP/s  (%.1f FMA/clk/SM @1.9GHz)\n",bps,ms,2*fma/ms/1e9, fma/(ms*1e-3)/sms/1.9e9);
    ms=timeit([&]{k_ffma2<16><<<sms*bps,256>>>(out,iters,1.0001f,0.9999f);});
    fma=(double)sms*bps*256*32*iters; printf("FFMA2 x16acc bps=%d: %.3f ms  %.2f TFLOP/s  (%.1f FMA/clk/SM @1.9GHz)\n",bps,ms,2*fma/ms/1e9, fma/(ms*1e-3)/sms/1.9e9);
  }
  int it2=200;
  for(int bps=1;bps<=2;bps++){
    float ms=timeit([&]{k_convlike<false><<<sms*bps,256>>>(out,it2);});
    double fma=(double)sms*bps*256*(double)it2*8*27*32; printf("convlike FFMA  bps=%d: %.3f ms %.2f TFLOP/s\n",bps,ms,2*fma/ms/1e9);
    ms=timeit([&]{k_convlike<true><<<sms*bps,256>>>(out,it2);});
    printf("convlike FFMA2 bps=%d: %.3f ms %.2f TFLOP/s\n",bps,ms,2*fma/ms/1e9);
  }
  return 0;
}
