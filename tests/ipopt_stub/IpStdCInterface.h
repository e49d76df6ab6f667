/* IpStdCInterface.h -- TEST STUB of IPOPT's C interface (types and the entry points src/eCUDA/ecuda_nlp_ipopt.cpp
 * uses, signatures as in IPOPT 3.11-3.14). IPOPT is not installed in this image; this header exists so that the
 * adapter is compiled and its callbacks are exercised by tests/test_ipopt_adapter.py instead of rotting behind an
 * #ifdef. The "solver" behind it (tests/ipopt_stub/stub_ipopt.cpp) is not an optimiser: it queries both structures,
 * evaluates every callback once at the starting point and returns Solve_Succeeded. */
#ifndef TESTS_IPOPT_STUB_IPSTDCINTERFACE_H_
#define TESTS_IPOPT_STUB_IPSTDCINTERFACE_H_
#ifdef __cplusplus
extern "C" {
#endif
typedef double Number;
typedef int Index;
typedef int Int;
typedef int Bool;
#ifndef TRUE
#define TRUE (1)
#endif
#ifndef FALSE
#define FALSE (0)
#endif
typedef void* UserDataPtr;
struct IpoptProblemInfo;
typedef struct IpoptProblemInfo* IpoptProblem;
typedef Bool (*Eval_F_CB)(Index n, Number* x, Bool new_x, Number* obj_value, UserDataPtr user_data);
typedef Bool (*Eval_Grad_F_CB)(Index n, Number* x, Bool new_x, Number* grad_f, UserDataPtr user_data);
typedef Bool (*Eval_G_CB)(Index n, Number* x, Bool new_x, Index m, Number* g, UserDataPtr user_data);
typedef Bool (*Eval_Jac_G_CB)(Index n, Number* x, Bool new_x, Index m, Index nele_jac, Index* iRow, Index* jCol,
                              Number* values, UserDataPtr user_data);
typedef Bool (*Eval_H_CB)(Index n, Number* x, Bool new_x, Number obj_factor, Index m, Number* lambda, Bool new_lambda,
                          Index nele_hess, Index* iRow, Index* jCol, Number* values, UserDataPtr user_data);
IpoptProblem CreateIpoptProblem(Index n, Number* x_L, Number* x_U, Index m, Number* g_L, Number* g_U, Index nele_jac,
                                Index nele_hess, Index index_style, Eval_F_CB eval_f, Eval_G_CB eval_g,
                                Eval_Grad_F_CB eval_grad_f, Eval_Jac_G_CB eval_jac_g, Eval_H_CB eval_h);
void FreeIpoptProblem(IpoptProblem ipopt_problem);
Bool AddIpoptStrOption(IpoptProblem ipopt_problem, char* keyword, char* val);
Bool AddIpoptNumOption(IpoptProblem ipopt_problem, char* keyword, Number val);
Bool AddIpoptIntOption(IpoptProblem ipopt_problem, char* keyword, Int val);
int IpoptSolve(IpoptProblem ipopt_problem, Number* x, Number* g, Number* obj_val, Number* mult_g, Number* mult_x_L,
               Number* mult_x_U, UserDataPtr user_data);
#ifdef __cplusplus
}
#endif
#endif
