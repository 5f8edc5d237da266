// Experiment: streaming map kernel variants (add, gelu) — which structure reaches the HBM roofline?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

__device__ __forceinline__ float4 ld_nc(const float4* p){ float4 v; asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.f32 {%0,%1,%2,%3}, [%4];":"=f"(v.x),"=f"(v.y),"=f"(v.z),"=f"(v.w):"l"(p)); return v;}
__device__ __forceinline__ float4 ld_plain(const float4* p){ return *p; }
__device__ __forceinline__ void st_cs(float4* p, float4 v){ asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};"::"l"(p),"f"(v.x),"f"(v.y),"f"(v.z),"f"(v.w):"memory"); }
__device__ __forceinline__ void st_plain(float4* p, float4 v){ *p = v; }

template<int OP> __device__ __forceinline__ float op1(float x, float y){
  if (OP==0) return x+y;
  if (OP==1) { // gelu current
    const float x3 = __fmul_rn(__fmul_rn(x, x), x);
    const float u = __fmul_rn(0.7978846f, __fadd_rn(x, __fmul_rn(0.044715f, x3)));
    return x / (1.0f + expf(-2.0f * u));
  }
  if (OP==2) { // gelu rcp + newton
    const float x3 = __fmul_rn(__fmul_rn(x, x), x);
    const float u = __fmul_rn(0.7978846f, __fadd_rn(x, __fmul_rn(0.044715f, x3)));
    const float d = 1.0f + expf(-2.0f * u);
    float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
    r = fmaf(r, fmaf(-d, r, 1.0f), r);
    return x * r;
  }
  if (OP==3) { // sigmoid current
    const float s = 1.0f / (1.0f + expf(-x));
    return x < -50.0f ? 0.0f : (x > 50.0f ? 1.0f : s);
  }
  return x;
}
template<int OP> __device__ __forceinline__ float4 op4(float4 a, float4 b){ return make_float4(op1<OP>(a.x,b.x),op1<OP>(a.y,b.y),op1<OP>(a.z,b.z),op1<OP>(a.w,b.w)); }

// V0: persistent, 4 loads in flight, load->compute->store (current design)
template<int OP, bool BIN, int U> __global__ void __launch_bounds__(256) k_persist(const float4* a, const float4* b, float4* o, size_t nvec){
  const size_t tile = 256*U; const size_t tiles = nvec / tile;
  for (size_t t = blockIdx.x; t < tiles; t += gridDim.x){
    size_t base = t*tile + threadIdx.x; float4 x[U], y[U];
    #pragma unroll
    for (int u=0;u<U;++u) x[u]=ld_nc(a+base+u*256);
    if (BIN){
    #pragma unroll
    for (int u=0;u<U;++u) y[u]=ld_nc(b+base+u*256);}
    #pragma unroll
    for (int u=0;u<U;++u) st_cs(o+base+u*256, op4<OP>(x[u], BIN?y[u]:x[u]));
  }
}
// V1: persistent with register prefetch of the next tile
template<int OP, bool BIN, int U> __global__ void __launch_bounds__(256) k_prefetch(const float4* a, const float4* b, float4* o, size_t nvec){
  const size_t tile = 256*U; const size_t tiles = nvec / tile;
  float4 x[U], y[U], nx[U], ny[U];
  size_t t = blockIdx.x;
  if (t < tiles){ size_t base=t*tile+threadIdx.x;
    #pragma unroll
    for (int u=0;u<U;++u){ nx[u]=ld_nc(a+base+u*256); if (BIN) ny[u]=ld_nc(b+base+u*256);} }
  for (; t < tiles; t += gridDim.x){
    #pragma unroll
    for (int u=0;u<U;++u){ x[u]=nx[u]; if (BIN) y[u]=ny[u]; }
    size_t tn = t + gridDim.x;
    if (tn < tiles){ size_t base=tn*tile+threadIdx.x;
      #pragma unroll
      for (int u=0;u<U;++u){ nx[u]=ld_nc(a+base+u*256); if (BIN) ny[u]=ld_nc(b+base+u*256);} }
    size_t base = t*tile + threadIdx.x;
    #pragma unroll
    for (int u=0;u<U;++u) st_cs(o+base+u*256, op4<OP>(x[u], BIN?y[u]:x[u]));
  }
}
// V2: non-persistent: one tile per block (grid = tiles), plain or hinted accesses
template<int OP, bool BIN, int U, bool HINT, int T> __global__ void __launch_bounds__(T) k_flat(const float4* a, const float4* b, float4* o, size_t nvec){
  size_t base = (size_t)blockIdx.x*T*U + threadIdx.x; float4 x[U], y[U];
  #pragma unroll
  for (int u=0;u<U;++u) x[u]= HINT? ld_nc(a+base+u*T) : ld_plain(a+base+u*T);
  if (BIN){
  #pragma unroll
  for (int u=0;u<U;++u) y[u]= HINT? ld_nc(b+base+u*T) : ld_plain(b+base+u*T);}
  #pragma unroll
  for (int u=0;u<U;++u){ float4 r=op4<OP>(x[u], BIN?y[u]:x[u]); if (HINT) st_cs(o+base+u*T, r); else st_plain(o+base+u*T, r);} 
}

template<class F> float timeit(F f, int iters=20){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for(int i=0;i<3;++i) f();
  std::vector<float> ts;
  for(int i=0;i<iters;++i){ cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); ts.push_back(ms);} 
  std::sort(ts.begin(), ts.end()); return ts[ts.size()/2];
}
int main(){
  size_t n = (size_t)4096*32000; size_t nvec=n/4;
  float *a,*b,*o; CK(cudaMalloc(&a,n*4)); CK(cudaMalloc(&b,n*4)); CK(cudaMalloc(&o,n*4));
  std::vector<float> h(n); for(size_t i=0;i<n;++i) h[i]=(float)((i*2654435761u)%2000)/250.f-4.f;
  CK(cudaMemcpy(a,h.data(),n*4,cudaMemcpyHostToDevice)); CK(cudaMemcpy(b,h.data(),n*4,cudaMemcpyHostToDevice));
  const float4 *a4=(const float4*)a,*b4=(const float4*)b; float4* o4=(float4*)o;
  auto rep=[&](const char* name, float ms, double bytes){ printf("%-44s %.4f ms  %.0f GB/s\n", name, ms, bytes/ms/1e6); };
  double b3=12.0*n, b2=8.0*n;
  int occ;
  #define PERSIST(K, name, bytes, ...) { cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, K<__VA_ARGS__>, 256, 0); for (int mult : {occ, 2*occ, 4*occ}) { int g=148*mult; char nm[128]; snprintf(nm,128,"%s grid=148x%d", name, mult); rep(nm, timeit([&]{ K<__VA_ARGS__><<<g,256>>>(a4,b4,o4,nvec); }), bytes);} }
  PERSIST(k_persist, "add persist U4", b3, 0,true,4)
  PERSIST(k_persist, "add persist U8", b3, 0,true,8)
  PERSIST(k_prefetch, "add prefetch U4", b3, 0,true,4)
  PERSIST(k_prefetch, "add prefetch U2", b3, 0,true,2)
  rep("add flat U4 T256 hint", timeit([&]{ k_flat<0,true,4,true,256><<<(unsigned)(nvec/(256*4)),256>>>(a4,b4,o4,nvec); }), b3);
  rep("add flat U4 T256 plain", timeit([&]{ k_flat<0,true,4,false,256><<<(unsigned)(nvec/(256*4)),256>>>(a4,b4,o4,nvec); }), b3);
  rep("add flat U2 T256 plain", timeit([&]{ k_flat<0,true,2,false,256><<<(unsigned)(nvec/(256*2)),256>>>(a4,b4,o4,nvec); }), b3);
  rep("add flat U1 T128 plain", timeit([&]{ k_flat<0,true,1,false,128><<<(unsigned)(nvec/(128)),128>>>(a4,b4,o4,nvec); }), b3);
  rep("add flat U4 T128 plain", timeit([&]{ k_flat<0,true,4,false,128><<<(unsigned)(nvec/(128*4)),128>>>(a4,b4,o4,nvec); }), b3);
  rep("add flat U8 T256 hint", timeit([&]{ k_flat<0,true,8,true,256><<<(unsigned)(nvec/(256*8)),256>>>(a4,b4,o4,nvec); }), b3);
  PERSIST(k_persist, "gelu persist U4", b2, 1,false,4)
  PERSIST(k_prefetch, "gelu prefetch U4", b2, 1,false,4)
  PERSIST(k_prefetch, "gelu(rcp) prefetch U4", b2, 2,false,4)
  PERSIST(k_persist, "gelu(rcp) persist U4", b2, 2,false,4)
  rep("gelu flat U4 T256 hint", timeit([&]{ k_flat<1,false,4,true,256><<<(unsigned)(nvec/(256*4)),256>>>(a4,b4,o4,nvec); }), b2);
  rep("gelu flat U4 T256 plain", timeit([&]{ k_flat<1,false,4,false,256><<<(unsigned)(nvec/(256*4)),256>>>(a4,b4,o4,nvec); }), b2);
  rep("gelu(rcp) flat U4 T256 plain", timeit([&]{ k_flat<2,false,4,false,256><<<(unsigned)(nvec/(256*4)),256>>>(a4,b4,o4,nvec); }), b2);
  rep("gelu(rcp) flat U2 T256 plain", timeit([&]{ k_flat<2,false,2,false,256><<<(unsigned)(nvec/(256*2)),256>>>(a4,b4,o4,nvec); }), b2);
  rep("gelu(rcp) flat U1 T256 plain", timeit([&]{ k_flat<2,false,1,false,256><<<(unsigned)(nvec/(256*1)),256>>>(a4,b4,o4,nvec); }), b2);
  rep("gelu(rcp) flat U2 T128 hint", timeit([&]{ k_flat<2,false,2,true,128><<<(unsigned)(nvec/(128*2)),128>>>(a4,b4,o4,nvec); }), b2);
  rep("sigmoid flat U4 T256 plain", timeit([&]{ k_flat<3,false,4,false,256><<<(unsigned)(nvec/(256*4)),256>>>(a4,b4,o4,nvec); }), b2);
  PERSIST(k_prefetch, "sigmoid prefetch U4", b2, 3,false,4)
  rep("copy flat U4 T256 plain", timeit([&]{ k_flat<4,false,4,false,256><<<(unsigned)(nvec/(256*4)),256>>>(a4,b4,o4,nvec); }), b2);
  PERSIST(k_persist, "copy persist U4", b2, 4,false,4)
  { float ms=timeit([&]{ cudaMemcpyAsync(o,a,n*4,cudaMemcpyDeviceToDevice); }); rep("cudaMemcpy D2D", ms, b2);} 
  CK(cudaDeviceSynchronize());
  return 0;
}
