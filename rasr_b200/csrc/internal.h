// internal.h -- cross-translation-unit accessors (not part of the C ABI)
#pragma once
#include "common.cuh"

struct rb_frontend;
struct rb_gmm;

cudaStream_t   rb_frontend_stream(const rb_frontend* h);
// pipeline.cu: drop the per-handle scratch buffers of the host-buffer pipelines (rb_frontend_destroy calls it)
void           rb_pipeline_forget(const rb_frontend* fe);
// host-buffer pipelines: upload the per-call tile tables on this (H2D) stream; nullptr: on the kernel stream
void           rb_frontend_set_upload_stream(rb_frontend* h, cudaStream_t s);
rb::DeviceInfo rb_frontend_device(const rb_frontend* h);
int            rb_frontend_feat_dim(const rb_frontend* h);
int            rb_frontend_convert_s16_dev(const rb_frontend* h, const int16_t* d_pcm, float* d_out, long n, int channels,
                                           int track, cudaStream_t s);
// gmm.cu: size the scorer's per-call scratch for calls of up to `frames` frames before a slab pipeline starts
int            rb_gmm_reserve(rb_gmm* h, long frames);
