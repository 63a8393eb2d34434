// dmp_stages.cu -- standalone stage kernels (a)-(e); placeholder until the fused path is verified.
#include "dmp_common.cuh"
extern "C" {
int dmp_stage_move(const DmpState*, const DmpIO*, int32_t*, void*) { return DMP_EINVAL; }
int dmp_stage_deposit(const DmpState*, const DmpIO*, int32_t*, void*) { return DMP_EINVAL; }
int dmp_stage_observe(const DmpState*, const DmpIO*, int32_t*, void*) { return DMP_EINVAL; }
int dmp_stage_reward(const DmpState*, const DmpIO*, int32_t*, void*) { return DMP_EINVAL; }
int dmp_stage_done_reset(const DmpState*, const DmpIO*, int32_t*, void*) { return DMP_EINVAL; }
}
