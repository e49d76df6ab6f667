// TrajectoryOptimizer.hpp -- the VGP container and eSolver plugin interface, source-compatible with
// the part of the reference's include/ETOL/TrajectoryOptimizer.hpp that lies on the VGP -> plugin
// path: the four pure virtuals (:39-54), loadConfigs / addParams /
// addExclZone / addAdjTrack / save, the setters and getters (:319-649) and the protected members
// (:664-693). Plotting, animation and the CGAL convex partition are visualisation / MIP-only and
// are not part of this tree.
#ifndef INCLUDE_ETOL_TRAJECTORYOPTIMIZER_HPP_
#define INCLUDE_ETOL_TRAJECTORYOPTIMIZER_HPP_

#include <algorithm>
#include <cfloat>
#include <iterator>
#include <list>
#include <string>
#include <vector>

#include <ETOL/ETOL_Types.hpp>

namespace ETOL {

class TrajectoryOptimizer {
 public:
    TrajectoryOptimizer();
    virtual ~TrajectoryOptimizer() {}

    // ---- the plugin interface every eSolver implements ------------------------------------------
    virtual void setup() = 0;  // build the solver's problem from the VGP
    virtual void solve() = 0;  // run it; on success setScore() + fill _xtraj/_utraj
    virtual void debug() = 0;  // raise verbosity (call after setup, before solve)
    virtual void close() = 0;  // release solver resources

    // ---- configuration ---------------------------------------------------------------------------
    const double getScore() const;
    void resetConfigs();
    void printConfigs();
    void loadConfigs(const char* filepath);   // ETOL XML: <etol><states><controls><exzones><mexzones>
    void saveConfigs(const char* filepath);
    void addParams(std::list<param_t> params);
    void addExclZone(border_t* border);  // keeps the raw border and its convex partition (getObstacles)
    void addAdjTrack(track_t* track);
    // convex partition of a simple polygon: per piece a lower and an upper chain, both sorted left to right
    // (reference TrajectoryOptimizer.hpp:118-128); slopes and lengths of the chain edges for the MIP formulations
    static region_t genRegion(border_t* border);
    static void calcSlopes(const region_t& region, std::vector<seg_t>* lowers, std::vector<seg_t>* uppers);

    // trajectory -> CSV. Never overwrites: bumps the trailing integer of the stem until the name is free.
    static std::string save(traj_t* traj, std::string fp = "traj.csv");

    // piecewise-linear lookup (reference TrajectoryOptimizer.hpp:239-257): below the table -> first
    // interval, above -> last interval, inside -> the last interval that brackets tval
    template <class T>
    static T linear_interpolation(const T& tval, const state_t& tvec, const state_t& ref) {
        size_t j = 0;
        if (tval > tvec.back()) {
            j = tvec.size() - 2;
        } else if (tval >= tvec.front()) {
            for (size_t c = 0; c + 1 < tvec.size(); ++c)
                if (tval >= tvec[c] && tval <= tvec[c + 1]) j = c;
        }
        return (tval - tvec.at(j)) * (ref.at(j + 1) - ref.at(j)) / (tvec.at(j + 1) - tvec.at(j)) + ref.at(j);
    }

    // columns of a trajectory: index 0 selects time, i>0 selects component i-1
    template <typename T>
    static traj_t extractTraj(const traj_t& traj, const std::vector<T>& idxs) {
        traj_t out;
        out.reserve(traj.size());
        for (const traj_elem_t& e : traj) {
            state_t s;
            for (const T& i : idxs) s.push_back(i == 0 ? e.first : e.second.at(i - 1));
            out.emplace_back(e.first, s);
        }
        return out;
    }
    template <typename T>
    static void scaleTraj(traj_t* traj, const std::vector<T>& scalers) {
        for (traj_elem_t& e : *traj)
            for (size_t i = 0; i < e.second.size(); ++i) e.second[i] *= (i < scalers.size()) ? scalers[i] : 1.;
    }
    template <typename T>
    static void offsetTraj(traj_t* traj, const std::vector<T>& offsets) {
        for (traj_elem_t& e : *traj)
            for (size_t i = 0; i < e.second.size(); ++i) e.second[i] += (i < offsets.size()) ? offsets[i] : 0.;
    }

    // ---- setters / getters -----------------------------------------------------------------------
    state_t& getX0();
    void setX0(const state_t& x0);
    state_t& getXf();
    void setXf(const state_t& xf);
    const size_t getNControls() const;
    const size_t getNStates() const;
    state_t& getXlower();
    void setXlower(const state_t& xlower);
    state_t& getXupper();
    void setXupper(const state_t& xupper);
    state_var_t& getXvartype();
    void setXvartype(const state_var_t& xvartype);
    const double getDt() const;
    void setDt(double dt);
    const size_t getNSteps() const;
    void setNSteps(const size_t nSteps);
    state_t& getXtol();
    void setXtol(const state_t& xtol);
    state_t& getUlower();
    void setUlower(const state_t& ulower);
    state_t& getUupper();
    void setUupper(const state_t& uupper);
    state_var_t& getUvartype();
    void setUvartype(const state_var_t& uvartype);
    const size_t getUrhorizon() const;
    void setUrhorizon(const size_t nu4dyn);
    const size_t getXrhorizon() const;
    void setXrhorizon(const size_t nx4dyn);
    size_t getRhorizon() const;
    void setNControls(const size_t nControls);
    void setNStates(const size_t nStates);
    void setEqConstraints(std::vector<f_t*> constraints);
    void setLessEqConstraints(std::vector<f_t*> constraints);
    void setConstraints(std::vector<f_t*> constraints);
    void setGradient(std::vector<f_t*> gradient);
    void setObjective(f_t* objective);
    traj_t* getUtraj();
    traj_t* getXtraj();
    const f_t* getObjective() const;
    std::vector<f_t*>* getGradient();
    std::vector<f_t*>* getEqConstraints();
    std::vector<f_t*>* getLessEqConstraints();
    std::vector<f_t*>* getConstraints();
    std::vector<border_t>* getObstacles_Raw();
    std::list<region_t>* getObstacles();
    std::list<track_t>* getTracks();
    paramset_t* getParams();  // addition: read access to the path-parameter map
    bool isMaximized() const;
    void setMaximize(const bool maximize);
    size_t getNExclZones();
    size_t getNTracks();

 protected:
    void errorHandler();
    void setScore(const double score);

    bool _maximize;
    double _score;
    double _dt;
    size_t _nSteps;
    size_t _nStates;
    size_t _nControls;
    state_t _x0;
    state_t _xlower;
    state_t _xupper;
    state_var_t _xvartype;
    state_t _xtol;
    state_t _xf;
    state_t _ulower;
    state_t _uupper;
    state_var_t _uvartype;
    size_t _xrhorizon;
    size_t _urhorizon;
    size_t _rhorizon;
    paramset_t _parameters;
    std::vector<border_t> _obstacles_raw;
    std::list<region_t> _obstacles;  // convex partitions (genRegion): what the MIP eSolvers of the reference consume
    std::list<track_t> _tracks;
    std::vector<f_t*> _constraints;
    std::vector<f_t*> _eq;
    std::vector<f_t*> _lesseq;
    std::vector<f_t*> _gradient;
    f_t* _objective;
    traj_t _xtraj;
    traj_t _utraj;
    std::bad_any_cast* _eAny;
};

}  // namespace ETOL
#endif  // INCLUDE_ETOL_TRAJECTORYOPTIMIZER_HPP_
