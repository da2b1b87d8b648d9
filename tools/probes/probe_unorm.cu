// Probe: is the texture unit's unorm8 -> float conversion (cudaReadModeNormalizedFloat) exactly fl(c / 255.0f)?
// nvcc -gencode arch=compute_100a,code=sm_100a -o probe_unorm probe_unorm.cu && ./probe_unorm
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
__global__ void k(cudaTextureObject_t t, float* out, float* outg) {
    const int i = threadIdx.x;
    const float4 v = tex2D<float4>(t, i + 0.5f, 0.5f);
    out[i] = v.x;
    const float4 g = tex2Dgather<float4>(t, (float)i + 1.0f, 1.0f, 1);   // texels (i, i+1) x (0, 1), component y
    outg[i] = g.w;                                                         // (i0, j0)
}
int main() {
    uchar4 h[2][257];
    for (int j = 0; j < 2; ++j) for (int i = 0; i < 257; ++i) { unsigned char c = (unsigned char)(i & 255); h[j][i] = make_uchar4(c, c, c, c); }
    cudaChannelFormatDesc fmt = cudaCreateChannelDesc<uchar4>();
    cudaArray_t arr; cudaMallocArray(&arr, &fmt, 257, 2, cudaArrayTextureGather);
    cudaMemcpy2DToArray(arr, 0, 0, h, 257 * 4, 257 * 4, 2, cudaMemcpyHostToDevice);
    cudaResourceDesc res; memset(&res, 0, sizeof(res)); res.resType = cudaResourceTypeArray; res.res.array.array = arr;
    cudaTextureDesc td; memset(&td, 0, sizeof(td));
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp; td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeNormalizedFloat;
    cudaTextureObject_t tex; cudaCreateTextureObject(&tex, &res, &td, nullptr);
    float *d, *dg; cudaMalloc(&d, 256 * 4); cudaMalloc(&dg, 256 * 4);
    k<<<1, 256>>>(tex, d, dg);
    float o[256], og[256]; cudaMemcpy(o, d, sizeof(o), cudaMemcpyDeviceToHost); cudaMemcpy(og, dg, sizeof(og), cudaMemcpyDeviceToHost);
    int bad = 0, badg = 0;
    for (int c = 0; c < 256; ++c) { const float want = (float)c / 255.0f; bad += o[c] != want; badg += og[c] != want; if (o[c] != want) printf("c=%d tex %.9g want %.9g\n", c, o[c], want); }
    printf("unorm8 -> float via tex2D: %d of 256 differ from fl(c/255); via tex2Dgather: %d differ (%s)\n", bad, badg, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
