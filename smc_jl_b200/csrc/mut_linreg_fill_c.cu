// regression sizes between the fully specialised ones (general kernel only): K = 19, 21, 22, 23
#include "mutate_kernel.cuh"

namespace smc {
void register_linreg_fill_c(std::vector<KernelEntry>& t)
{
    t.push_back(LINREG_LITE(19));
    t.push_back(LINREG_LITE(21));
    t.push_back(LINREG_LITE(22));
    t.push_back(LINREG_LITE(23));
}
}  // namespace smc
