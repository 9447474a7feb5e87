#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s -> %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)
int main(){
  const size_t n = 1<<19, rows = 13, colb = n*8;
  double *h, *d, *h2, *d2;
  CK(cudaMallocHost(&h, rows*colb)); CK(cudaMalloc(&d, rows*colb));
  CK(cudaMallocHost(&h2, 4*colb)); CK(cudaMalloc(&d2, 4*colb));
  cudaStream_t s1, s2; cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto timeit = [&](const char* name, auto f){
    f(); cudaDeviceSynchronize();
    float best=1e9;
    for(int r=0;r<5;++r){ cudaEventRecord(e0,s1); f(); cudaEventRecord(e1,s1); cudaEventSynchronize(e1); cudaDeviceSynchronize(); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
    printf("%-50s %.3f ms  %.1f GB/s\n", name, best, rows*colb/best/1e6);
  };
  timeit("1D 54.5MB H2D", [&]{ cudaMemcpyAsync(d,h,rows*colb,cudaMemcpyHostToDevice,s1); });
  timeit("13 x 1D 4MB H2D", [&]{ for(size_t r=0;r<rows;++r) cudaMemcpyAsync(d+r*n,h+r*n,colb,cudaMemcpyHostToDevice,s1); });
  timeit("2D 13 rows x 4MB (contiguous) H2D", [&]{ cudaMemcpy2DAsync(d,colb,h,colb,colb,rows,cudaMemcpyHostToDevice,s1); });
  timeit("4 chunks of 2D 13 rows x 1MB H2D", [&]{ for(int c=0;c<4;++c) cudaMemcpy2DAsync(d+c*(n/4),colb,h+c*(n/4),colb,colb/4,rows,cudaMemcpyHostToDevice,s1); });
  timeit("4 chunks x 13 x 1D 1MB H2D", [&]{ for(int c=0;c<4;++c) for(size_t r=0;r<rows;++r) cudaMemcpyAsync(d+r*n+c*(n/4),h+r*n+c*(n/4),colb/4,cudaMemcpyHostToDevice,s1); });
  timeit("8 chunks of 2D 13 rows x 512KB H2D", [&]{ for(int c=0;c<8;++c) cudaMemcpy2DAsync(d+c*(n/8),colb,h+c*(n/8),colb,colb/8,rows,cudaMemcpyHostToDevice,s1); });
  timeit("1D 54.5MB H2D with concurrent 16MB D2H", [&]{ cudaMemcpyAsync(h2,d2,4*colb,cudaMemcpyDeviceToHost,s2); cudaMemcpyAsync(d,h,rows*colb,cudaMemcpyHostToDevice,s1); });
  timeit("4x2D H2D with concurrent 4x2D D2H", [&]{ for(int c=0;c<4;++c){ cudaMemcpy2DAsync(h2+c*(n/4),colb,d2+c*(n/4),colb,colb/4,4,cudaMemcpyDeviceToHost,s2); cudaMemcpy2DAsync(d+c*(n/4),colb,h+c*(n/4),colb,colb/4,rows,cudaMemcpyHostToDevice,s1);} });
  return 0;
}
