// regression sizes between the fully specialised ones (general kernel only): K = 25, 26, 27
#include "mutate_kernel.cuh"

namespace smc {
void register_linreg_fill_d(std::vector<KernelEntry>& t)
{
    t.push_back(LINREG_LITE(25));
    t.push_back(LINREG_LITE(26));
    t.push_back(LINREG_LITE(27));
}
}  // namespace smc
