// TEST INFRASTRUCTURE ONLY (oracle/_ref build).  Not part of the product.
//
// Host launchers for the spliced reference kernels.  The launch geometry restates
// ImResizeLayer::Forward_gpu (imresize_layer.cu:157-190: one launch per channel, 16x16 blocks)
// and NmsLayer::Forward_gpu (nms_layer.cu:116-182: per part  register -> thrust::exclusive_scan
// -> writeResult, 256-thread blocks).  Host-pointer entry points so ctypes can call them.
#pragma once

#define REF_CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "ref_cpm: %s -> %s\n", #x, cudaGetErrorString(e_)); return -1; } } while (0)

extern "C" int ref_imresize_host(const float* src, float* dst, int num, int channel, int oh8, int ow8,
                                 int th, int tw, float start_scale, float scale_gap) {
    float *d_src = nullptr, *d_dst = nullptr;
    size_t nsrc = (size_t)num * channel * oh8 * ow8, ndst = (size_t)channel * th * tw;
    REF_CK(cudaMalloc(&d_src, nsrc * sizeof(float)));
    REF_CK(cudaMalloc(&d_dst, ndst * sizeof(float)));
    REF_CK(cudaMemcpy(d_src, src, nsrc * sizeof(float), cudaMemcpyHostToDevice));
    REF_CK(cudaMemset(d_dst, 0, ndst * sizeof(float)));
    const dim3 threadsPerBlock(NUMBER_THREADS_PER_BLOCK_1D, NUMBER_THREADS_PER_BLOCK_1D);
    const dim3 numBlocks(caffe::updiv(tw, threadsPerBlock.x), caffe::updiv(th, threadsPerBlock.y));
    const int offset_src = oh8 * ow8, offset_dst = tw * th;
    for (int c = 0; c < channel; c++) {
        caffe::imresize_cubic_kernel<float><<<numBlocks, threadsPerBlock>>>(
            d_src + c * offset_src, d_dst + c * offset_dst, channel * offset_src, num, scale_gap, start_scale,
            ow8, oh8, tw, th);
    }
    REF_CK(cudaGetLastError());
    REF_CK(cudaDeviceSynchronize());
    REF_CK(cudaMemcpy(dst, d_dst, ndst * sizeof(float), cudaMemcpyDeviceToHost));
    cudaFree(d_src); cudaFree(d_dst);
    return 0;
}

// src: `channels` full-resolution maps (channels >= num_parts + 1 so that the width-for-height
// window test, nms_layer.cu:79, aliases into a real following channel as it does in the net).
extern "C" int ref_nms_host(const float* src, float* dst, int channels, int height, int width,
                            int num_parts, int max_peaks, float threshold) {
    float *d_src = nullptr, *d_dst = nullptr; int* d_ws = nullptr;
    const int offset = height * width, offset_dst = (max_peaks + 1) * 3;
    REF_CK(cudaMalloc(&d_src, (size_t)channels * offset * sizeof(float)));
    REF_CK(cudaMalloc(&d_ws, (size_t)num_parts * offset * sizeof(int)));
    REF_CK(cudaMalloc(&d_dst, (size_t)num_parts * offset_dst * sizeof(float)));
    REF_CK(cudaMemcpy(d_src, src, (size_t)channels * offset * sizeof(float), cudaMemcpyHostToDevice));
    REF_CK(cudaMemset(d_dst, 0, (size_t)num_parts * offset_dst * sizeof(float)));
    REF_CK(cudaMemset(d_ws, 0, (size_t)num_parts * offset * sizeof(int)));
    const dim3 threadsPerBlock(NUMBER_THREADS_PER_BLOCK_1D, NUMBER_THREADS_PER_BLOCK_1D);
    const dim3 numBlocks(caffe::updiv(width, threadsPerBlock.x), caffe::updiv(height, threadsPerBlock.y));
    for (int c = 0; c < num_parts; c++) {
        int* w_pointer1 = d_ws + c * offset;
        const float* s = d_src + c * offset;
        float* d = d_dst + c * offset_dst;
        caffe::nms_register_kernel<float><<<numBlocks, threadsPerBlock>>>(s, w_pointer1, width, height, threshold);
        thrust::device_ptr<int> dev_ptr = thrust::device_pointer_cast(w_pointer1);
        thrust::exclusive_scan(dev_ptr, dev_ptr + offset, dev_ptr);
        caffe::writeResultKernel<float><<<caffe::updiv(offset, NUMBER_THREADS_PER_BLOCK), NUMBER_THREADS_PER_BLOCK>>>(
            offset, w_pointer1, s, d, width, max_peaks);
    }
    REF_CK(cudaGetLastError());
    REF_CK(cudaDeviceSynchronize());
    REF_CK(cudaMemcpy(dst, d_dst, (size_t)num_parts * offset_dst * sizeof(float), cudaMemcpyDeviceToHost));
    cudaFree(d_src); cudaFree(d_dst); cudaFree(d_ws);
    return 0;
}
