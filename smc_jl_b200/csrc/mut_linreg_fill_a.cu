// regression sizes between the fully specialised ones (general kernel only): K = 7, 9, 11, 13
#include "mutate_kernel.cuh"

namespace smc {
void register_linreg_fill_a(std::vector<KernelEntry>& t)
{
    t.push_back(LINREG_LITE(7));
    t.push_back(LINREG_LITE(9));
    t.push_back(LINREG_LITE(11));
    t.push_back(LINREG_LITE(13));
}
}  // namespace smc
