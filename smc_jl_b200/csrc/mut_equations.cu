// mut_equations.cu -- instantiates the mutation / evaluation kernels of these likelihood functors (see mutate_kernel.cuh)
#include "mutate_kernel.cuh"

namespace smc {
void register_equations(std::vector<KernelEntry>& t)
{
    t.push_back(make_entry<GaussReg<3, 2, 3, 0, 2>>());   /* test/modelsetup.jl 3-equation model, CAPM (per-period form) */
    t.push_back(make_entry<GaussReg<3, 1, 3, 0, 2>>());   /* examples/capm_model as written */
    t.push_back(make_entry<GaussReg<1, 2, 3, 0, 2>>());   /* one equation (alpha, beta, sigma) */
    t.push_back(make_entry<GaussReg<2, 2, 3, 0, 2>>());
}
}  // namespace smc
