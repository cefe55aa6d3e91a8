// regression sizes between the fully specialised ones (general kernel only): K = 14, 15, 17, 18
#include "mutate_kernel.cuh"

namespace smc {
void register_linreg_fill_b(std::vector<KernelEntry>& t)
{
    t.push_back(LINREG_LITE(14));
    t.push_back(LINREG_LITE(15));
    t.push_back(LINREG_LITE(17));
    t.push_back(LINREG_LITE(18));
}
}  // namespace smc
