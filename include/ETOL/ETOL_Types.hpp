// ETOL_Types.hpp -- vocabulary types of the VGP container, source-compatible with the reference's
// include/ETOL/ETOL_Types.hpp (same names, same shapes) so that code written against ETOL compiles
// unchanged against this tree. Fresh implementation: only std headers, `using` aliases.
//   callback ABI   reference ETOL_Types.hpp:25-27,111-117 (F_ARGS, scalar_t, vector_t, f_t)
//   parameters     reference ETOL_Types.hpp:31-51
//   geometry       reference ETOL_Types.hpp:53-110
#ifndef INCLUDE_ETOL_ETOL_TYPES_HPP_
#define INCLUDE_ETOL_ETOL_TYPES_HPP_

#include <any>
#include <array>
#include <cmath>
#include <functional>
#include <list>
#include <map>
#include <string>
#include <utility>
#include <vector>

#define PARAM_PAIR ETOL::param_name_t, ETOL::param_configs_t
// every VGP callback has this signature; what the std::any elements hold is fixed per eSolver
// (adouble* for ePSOPT, double for eDymos, ecuda scalar pointers for eCUDA -- see eCUDA_Types.hpp)
#define F_ARGS ETOL::vector_t x, ETOL::vector_t u, ETOL::vector_t params, \
    std::vector<std::string> pnames, std::any k, std::any dt

namespace ETOL {

enum var_t { CONTINUOUS = 0, INTERGER = 1, BINARY = 2 };  // spelling of the reference kept

using param_name_t = std::string;
struct param_configs_t {
    var_t varType = var_t::CONTINUOUS;
    double lbnd = 0.;
    double ubnd = 0.;
    double tStart = 0.;  // time the parameter becomes active
    double tStop = 0.;   // time it stops being active
};
using param_t = std::pair<PARAM_PAIR>;
using paramset_t = std::map<PARAM_PAIR>;  // iterated in name order (ePSOPT.cpp:147-150 relies on it)

using coord_t = std::pair<double, double>;
using line_t = std::vector<coord_t>;
using lines_t = std::vector<line_t>;
using corner_t = std::array<double, 3>;
struct edge_prop_ {
    double slope = 0.;
    double length = 0.;
};
using edge_prop_t = edge_prop_;
using edge_t = std::pair<corner_t, edge_prop_t>;
using seg_t = std::vector<edge_t>;
using closure_t = std::pair<std::vector<seg_t>, std::vector<seg_t>>;
using border_t = std::list<corner_t>;  // polygon corners in order, closed implicitly

using state_t = std::vector<double>;
using state_var_t = std::vector<var_t>;
using traj_elem_t = std::pair<double, state_t>;  // (time, values)
using traj_t = std::vector<traj_elem_t>;

struct boundary_t {
    border_t lower;
    border_t upper;
};
struct track_t {  // a moving exclusion zone: keep-out radius around a waypoint trajectory
    double radius = 0.0;
    traj_t trajectory = traj_t();
};
using region_t = std::list<boundary_t>;

using scalar_t = std::any;
using vector_t = std::vector<scalar_t>;
using f_t = std::function<scalar_t(F_ARGS)>;

}  // namespace ETOL
#endif  // INCLUDE_ETOL_ETOL_TYPES_HPP_
