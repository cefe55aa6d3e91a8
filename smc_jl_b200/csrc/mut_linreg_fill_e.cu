// regression sizes between the fully specialised ones (general kernel only): K = 28, 29
#include "mutate_kernel.cuh"

namespace smc {
void register_linreg_fill_e(std::vector<KernelEntry>& t)
{
    t.push_back(LINREG_LITE(28));
    t.push_back(LINREG_LITE(29));
}
}  // namespace smc
