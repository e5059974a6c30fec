// TEST INFRASTRUCTURE ONLY (oracle/_ref build).  Not part of the product.
//
// Host entry point around the spliced reference render kernels.  Restates the dispatch of render() (examples/rtpose/
// rtpose.cpp:271-300) and of the three launchers (renderFunctions.cu:331-389, 978-1036, 1038-1080), including their
// swapped launch configuration `<<<threadsPerBlock, numBlocks>>>` (grid 32x32, block = (ceil(w/32), ceil(h/32))).
#pragma once

#define REFR_CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "ref_render: %s -> %s\n", #x, cudaGetErrorString(e_)); return -1; } } while (0)

// canvas: 3 x h_canvas x w_canvas float planar (in/out, host); heatmaps: num_maps x h_net x w_net (host, may be NULL
// when part_to_show == 0); poses: num_people x num_parts x 3 (host).
extern "C" int ref_render_host(float* canvas, int w_canvas, int h_canvas, int w_net, int h_net, const float* heatmaps,
                               int num_maps, const float* poses, int num_people, int num_parts, int part_to_show,
                               int googly_eyes) {
    float *d_canvas = nullptr, *d_heat = nullptr, *d_poses = nullptr;
    const size_t nc = (size_t)3 * w_canvas * h_canvas, nh = (size_t)num_maps * w_net * h_net;
    REFR_CK(cudaMalloc(&d_canvas, nc * sizeof(float)));
    REFR_CK(cudaMemcpy(d_canvas, canvas, nc * sizeof(float), cudaMemcpyHostToDevice));
    if (heatmaps) {
        REFR_CK(cudaMalloc(&d_heat, nh * sizeof(float)));
        REFR_CK(cudaMemcpy(d_heat, heatmaps, nh * sizeof(float), cudaMemcpyHostToDevice));
    }
    REFR_CK(cudaMalloc(&d_poses, (size_t)RENDER_MAX_PEOPLE * 70 * 3 * sizeof(float)));
    REFR_CK(cudaMemset(d_poses, 0, (size_t)RENDER_MAX_PEOPLE * 70 * 3 * sizeof(float)));
    if (num_people > 0) REFR_CK(cudaMemcpy(d_poses, poses, (size_t)num_people * num_parts * 3 * sizeof(float), cudaMemcpyHostToDevice));
    const dim3 threadsPerBlock(numThreadsPerBlock_1d, numThreadsPerBlock_1d);
    const dim3 numBlocks(caffe::updiv(w_canvas, threadsPerBlock.x), caffe::updiv(h_canvas, threadsPerBlock.y));
    const float ratio_to_origin = (float)h_canvas / (float)h_net;
    const int boxsize = 368;   // BOX_SIZE, rtpose.cpp:92 (unused by the kernels)
    if (num_parts == 15) {                       // render_mpi_parts
        if (part_to_show == 0) {
            if (num_people != 0)
                render_pose_29parts<<<threadsPerBlock, numBlocks>>>(d_canvas, w_canvas, h_canvas, ratio_to_origin, d_poses, boxsize, num_people, 0.0f);
        } else {
            render_pose_29parts_heatmap<<<threadsPerBlock, numBlocks>>>(d_canvas, w_canvas, h_canvas, w_net, h_net, d_heat, num_people, part_to_show - 1);
        }
    } else if (part_to_show - 1 <= num_parts) {   // render_coco_parts
        if (part_to_show == 0) {
            if (num_people != 0)
                render_pose_coco_parts<<<threadsPerBlock, numBlocks>>>(d_canvas, w_canvas, h_canvas, ratio_to_origin, d_poses, boxsize, num_people, 0.01f, googly_eyes != 0);
        } else if (part_to_show - 1 == num_parts) {
            render_pose_coco_heatmap2<<<threadsPerBlock, numBlocks>>>(d_canvas, w_canvas, h_canvas, w_net, h_net, d_heat, num_people, 0);
        } else {
            render_pose_coco_heatmap<<<threadsPerBlock, numBlocks>>>(d_canvas, w_canvas, h_canvas, w_net, h_net, d_heat, num_people, part_to_show - 1);
        }
    } else {                                      // render_coco_aff
        int aff_part = ((part_to_show - 1) - num_parts - 1) * 2;
        int num_parts_accum = 1;
        if (aff_part == 0) num_parts_accum = 19; else aff_part = aff_part - 2;
        aff_part += 1 + num_parts;
        render_pose_coco_affinity<<<threadsPerBlock, numBlocks>>>(d_canvas, w_canvas, h_canvas, w_net, h_net, d_heat, num_parts_accum, num_people, aff_part);
    }
    REFR_CK(cudaGetLastError());
    REFR_CK(cudaDeviceSynchronize());
    REFR_CK(cudaMemcpy(canvas, d_canvas, nc * sizeof(float), cudaMemcpyDeviceToHost));
    cudaFree(d_canvas); cudaFree(d_heat); cudaFree(d_poses);
    return 0;
}
