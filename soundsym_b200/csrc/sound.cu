// sound.cu — TEMPORARY stubs (replaced by the MFCC / segmentation kernels)
#include "sound.cuh"
using namespace ss;
extern "C" {
#define NOTIMPL(ctx) return set_error(ctx, SS_ERR_INVALID, "%s: not implemented yet", __func__)
int ss_decode_pcm(ss_ctx* ctx, const int32_t*, size_t, int, double*) { NOTIMPL(ctx); }
int ss_sound_analyze(ss_ctx* ctx, const double*, size_t, double, int, double*, size_t*, double*, double*) { NOTIMPL(ctx); }
int ss_mfcc(ss_ctx* ctx, const double*, size_t, double, int, double*, size_t*) { NOTIMPL(ctx); }
int ss_max_power(ss_ctx* ctx, const double*, size_t, double*) { NOTIMPL(ctx); }
int ss_mfcc_dev(ss_ctx* ctx, const double*, size_t, double, int, double*) { NOTIMPL(ctx); }
int ss_symbols(ss_ctx* ctx, const double*, size_t, const ss_gmm*, uint8_t*, double*) { NOTIMPL(ctx); }
int ss_vote_split(ss_ctx* ctx, const uint8_t*, size_t, int, int, uint32_t*, uint64_t*, size_t*) { NOTIMPL(ctx); }
int ss_partition(ss_ctx* ctx, const double*, size_t, const ss_gmm*, int, int, uint64_t*, size_t*) { NOTIMPL(ctx); }
int ss_resynth(ss_ctx* ctx, const double*, const uint64_t*, size_t, const uint32_t*, const uint64_t*, size_t, double*) { NOTIMPL(ctx); }
int ss_sequence_distances(ss_ctx* ctx, const double*, size_t, int, double*) { NOTIMPL(ctx); }
}
