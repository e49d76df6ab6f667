// TrajectoryOptimizer.cpp -- core-lite: the part of ETOL's VGP container that lies on the
// VGP -> eSolver path, written fresh for this tree (the reference core needs libxml2, CGAL, Boost and
// gnuplot-iostream, none of which the hot path uses).
//
// Behaviour follows the reference's src/TrajectoryOptimizer/TrajectoryOptimizer.cpp:
//   loadConfigs  :787-1117   XML wire format (<etol><states><controls><exzones><mexzones>)
//   save         :626-674    trajectory -> CSV, never overwriting
//   addParams / addExclZone / addAdjTrack  :1636-1653
//   setters / getters        :1655-1873
// Deviations (documented in DESIGN.md section 6):
//   * the parser is an expat SAX walk instead of a libxml2 DOM walk; numbers are read with strtod,
//     so exponents ("1e-3") are accepted where XPath 1.0 numbers would give NaN;
//   * resetConfigs() also clears _xlower/_xupper/_parameters/_nSteps, so loading twice into one
//     object replaces the VGP instead of appending to it;
//   * addExclZone partitions the border into convex pieces like the reference (:84-159), but without CGAL: ear
//     clipping followed by a Hertel-Mehlhorn merge instead of CGAL::optimal_convex_partition_2 -- a valid convex
//     partition with lower / upper chains per piece, not necessarily the minimum number of pieces. Only the MIP
//     eSolvers and plotting consume it; neither is part of this tree.
#include <ETOL/TrajectoryOptimizer.hpp>

#include <expat.h>
#include <sys/stat.h>

#include <algorithm>
#include <array>
#include <cassert>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>

namespace ETOL {

TrajectoryOptimizer::TrajectoryOptimizer()
    : _maximize(false), _score(0.), _dt(0.), _nSteps(0), _nStates(0), _nControls(0), _xrhorizon(0), _urhorizon(0),
      _rhorizon(0), _objective(nullptr), _eAny(nullptr) {}

// ---- XML --------------------------------------------------------------------------------------------------
namespace {

// SAX state: which element we are inside and the caps announced by the n* attributes
struct Loader {
    TrajectoryOptimizer* t = nullptr;
    std::vector<std::string> stack;
    bool seen_root = false;
    size_t nstates = 0, ncontrols = 0;
    size_t nzones = SIZE_MAX, ncorners = SIZE_MAX, zones_added = 0;
    size_t ntracks = SIZE_MAX, nwaypoints = SIZE_MAX, ndatums = SIZE_MAX;
    border_t border;
    track_t track;
    traj_elem_t waypoint;
    std::string text;
    bool in_datum = false;
    std::string error;
};

double number(const char* s) {
    char* end = nullptr;
    double v = std::strtod(s, &end);
    if (end == s) return std::nan("");
    return v;
}

var_t vartype(const char* s, Loader* L, const char* what) {
    switch (s[0]) {
        case 'C': return var_t::CONTINUOUS;
        case 'B': return var_t::BINARY;
        case 'I': return var_t::INTERGER;
        default:
            if (L->error.empty()) L->error = std::string("Invalid ") + what;
            return var_t::CONTINUOUS;
    }
}

const char* attr(const char** atts, const char* name) {
    for (int i = 0; atts[i]; i += 2)
        if (std::strcmp(atts[i], name) == 0) return atts[i + 1];
    return nullptr;
}

bool under(const Loader* L, const char* parent) { return !L->stack.empty() && L->stack.back() == parent; }

void XMLCALL on_start(void* ud, const char* name, const char** atts) {
    Loader* L = static_cast<Loader*>(ud);
    TrajectoryOptimizer* t = L->t;
    const std::string el(name);
    if (!L->seen_root && el == "etol") {  // the reference looks the root up by XPath //etol
        L->seen_root = true;
        if (const char* v = attr(atts, "nsteps")) t->setNSteps(static_cast<size_t>(number(v)));
        if (const char* v = attr(atts, "dt")) t->setDt(number(v));
    } else if (under(L, "etol") && el == "states") {
        L->nstates = 0;
        if (const char* v = attr(atts, "nstates")) L->nstates = static_cast<size_t>(number(v));
        if (const char* v = attr(atts, "rhorizon"))
            t->setXrhorizon(std::max(t->getXrhorizon(), static_cast<size_t>(number(v))));
    } else if (under(L, "states")) {
        // nstates is a cap: children are read while fewer than nstates have been counted, and a
        // child counts when its vartype attribute is seen. Every other attribute is appended to its
        // own vector independently (a missing attribute silently shortens that vector).
        if (L->nstates > t->getNStates()) {
            for (int i = 0; atts[i]; i += 2) {
                const char *a = atts[i], *v = atts[i + 1];
                if (!std::strcmp(a, "vartype")) {
                    t->setNStates(t->getNStates() + 1);
                    t->getXvartype().push_back(vartype(v, L, "xVartype"));
                } else if (!std::strcmp(a, "lower")) {
                    t->getXlower().push_back(number(v));
                } else if (!std::strcmp(a, "upper")) {
                    t->getXupper().push_back(number(v));
                } else if (!std::strcmp(a, "initial")) {
                    t->getX0().push_back(number(v));
                } else if (!std::strcmp(a, "terminal")) {
                    t->getXf().push_back(number(v));
                } else if (!std::strcmp(a, "tolerance")) {
                    t->getXtol().push_back(number(v));
                }
            }
        }
    } else if (under(L, "etol") && el == "controls") {
        L->ncontrols = 0;
        if (const char* v = attr(atts, "ncontrols")) L->ncontrols = static_cast<size_t>(number(v));
        if (const char* v = attr(atts, "rhorizon"))
            t->setUrhorizon(std::max(t->getUrhorizon(), static_cast<size_t>(number(v))));
    } else if (under(L, "controls")) {
        if (L->ncontrols > t->getNControls()) {
            for (int i = 0; atts[i]; i += 2) {
                const char *a = atts[i], *v = atts[i + 1];
                if (!std::strcmp(a, "vartype")) {
                    t->setNControls(t->getNControls() + 1);
                    t->getUvartype().push_back(vartype(v, L, "uVartype"));
                } else if (!std::strcmp(a, "lower")) {
                    t->getUlower().push_back(number(v));
                } else if (!std::strcmp(a, "upper")) {
                    t->getUupper().push_back(number(v));
                }
            }
        }
    } else if (under(L, "etol") && el == "exzones") {
        L->nzones = SIZE_MAX;
        L->zones_added = 0;
        if (const char* v = attr(atts, "nzones")) L->nzones = static_cast<size_t>(number(v));
    } else if (under(L, "exzones")) {  // <border ncorners=..>
        L->border.clear();
        L->ncorners = SIZE_MAX;
        if (const char* v = attr(atts, "ncorners")) L->ncorners = static_cast<size_t>(number(v));
    } else if (L->stack.size() >= 2 && L->stack[L->stack.size() - 2] == "exzones") {  // <corner x y z/>
        const char *x = attr(atts, "x"), *y = attr(atts, "y"), *z = attr(atts, "z");
        // a corner needs all three coordinates; the reference's cap test is !(size > ncorners), i.e.
        // up to ncorners + 1 corners are accepted -- kept, shipped files never reach it
        if (x && y && z && !(L->border.size() > L->ncorners)) L->border.push_back({number(x), number(y), number(z)});
    } else if (under(L, "etol") && el == "mexzones") {
        L->ntracks = SIZE_MAX;
        if (const char* v = attr(atts, "nzones")) L->ntracks = static_cast<size_t>(number(v));
    } else if (under(L, "mexzones")) {  // <track radius nwaypoints>
        L->track = track_t();
        L->nwaypoints = SIZE_MAX;
        if (const char* v = attr(atts, "radius")) L->track.radius = number(v);
        if (const char* v = attr(atts, "nwaypoints")) L->nwaypoints = static_cast<size_t>(number(v));
    } else if (L->stack.size() >= 2 && L->stack[L->stack.size() - 2] == "mexzones") {  // <waypoint t ndatums>
        L->waypoint = traj_elem_t();
        L->ndatums = SIZE_MAX;
        if (const char* v = attr(atts, "t")) L->waypoint.first = number(v);
        if (const char* v = attr(atts, "ndatums")) L->ndatums = static_cast<size_t>(number(v));
    } else if (L->stack.size() >= 3 && L->stack[L->stack.size() - 3] == "mexzones" && el == "datum") {
        L->in_datum = true;  // <datum>text</datum>
        L->text.clear();
    }
    L->stack.push_back(el);
}

void XMLCALL on_text(void* ud, const char* s, int len) {
    Loader* L = static_cast<Loader*>(ud);
    if (L->in_datum) L->text.append(s, len);
}

void XMLCALL on_end(void* ud, const char*) {
    Loader* L = static_cast<Loader*>(ud);
    TrajectoryOptimizer* t = L->t;
    const size_t depth = L->stack.size();
    L->stack.pop_back();
    if (depth >= 2 && L->stack.back() == "exzones") {  // </border>
        if (!L->border.empty() && L->nzones > L->zones_added) {
            t->addExclZone(&L->border);
            ++L->zones_added;
        }
    } else if (depth >= 4 && L->stack[depth - 4] == "mexzones" && L->in_datum) {  // </datum>
        L->in_datum = false;
        // ndatums = 0 means "no cap" in the reference (:1088-1089)
        if (!(L->ndatums != 0 && L->waypoint.second.size() >= L->ndatums))
            L->waypoint.second.push_back(number(L->text.c_str()));
    } else if (depth >= 3 && L->stack[depth - 3] == "mexzones") {  // </waypoint>
        if (!L->waypoint.second.empty() && L->nwaypoints > L->track.trajectory.size())
            L->track.trajectory.push_back(L->waypoint);
    } else if (depth >= 2 && L->stack.back() == "mexzones") {  // </track>
        if (!L->track.trajectory.empty() && L->ntracks > t->getNTracks()) t->addAdjTrack(&L->track);
    }
}

}  // namespace

void TrajectoryOptimizer::loadConfigs(const char* filepath) {
    resetConfigs();
    std::ifstream in(filepath, std::ios::binary);
    if (!in) {
        std::cerr << "Document not parsed successfully: " << filepath << std::endl;
        exit(EXIT_FAILURE);
    }
    std::string doc((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    Loader L;
    L.t = this;
    XML_Parser parser = XML_ParserCreate(nullptr);
    XML_SetUserData(parser, &L);
    XML_SetElementHandler(parser, on_start, on_end);
    XML_SetCharacterDataHandler(parser, on_text);
    const bool ok = XML_Parse(parser, doc.data(), static_cast<int>(doc.size()), 1) != XML_STATUS_ERROR;
    if (!ok) {
        std::cerr << "Document not parsed successfully: " << filepath << ": "
                  << XML_ErrorString(XML_GetErrorCode(parser)) << " at line " << XML_GetCurrentLineNumber(parser)
                  << std::endl;
        XML_ParserFree(parser);
        exit(EXIT_FAILURE);
    }
    XML_ParserFree(parser);
    if (!L.error.empty()) {
        std::cout << L.error << std::endl;
        exit(EXIT_FAILURE);
    }
    if (!L.seen_root || getNSteps() == 0 || getDt() == 0) {
        std::cerr << "ETOL config needs an <etol> root with non-zero nsteps and dt: " << filepath << std::endl;
        exit(EXIT_FAILURE);
    }
}

// XML writer: same elements and attribute names as the files loadConfigs reads, numbers with
// two decimals like the reference writer (TrajectoryOptimizer.cpp:1119-1635, "%.2f"). Unlike the
// reference it writes the raw borders (there is no convex partition here) and loops controls over
// getNControls().
void TrajectoryOptimizer::saveConfigs(const char* filepath) {
    FILE* f = std::fopen(filepath, "w");
    if (!f) {
        std::cerr << "cannot write " << filepath << std::endl;
        return;
    }
    auto vt = [](var_t v) { return v == var_t::BINARY ? 'B' : (v == var_t::INTERGER ? 'I' : 'C'); };
    auto at = [](const state_t& v, size_t i) { return i < v.size() ? v[i] : 0.0; };
    std::fprintf(f, "<?xml version=\"1.0\" encoding=\"UTF-8\"?>\n<etol nsteps=\"%zu\" dt=\"%.2f\">\n", _nSteps, _dt);
    std::fprintf(f, "\t<states nstates=\"%zu\" rhorizon=\"%zu\">\n", _nStates, _xrhorizon);
    for (size_t i = 0; i < _nStates; ++i)
        std::fprintf(f,
                     "\t\t<state name=\"x%zu\" vartype=\"%c\" lower=\"%.2f\" upper=\"%.2f\" initial=\"%.2f\" "
                     "terminal=\"%.2f\" tolerance=\"%.2f\"/>\n",
                     i, vt(i < _xvartype.size() ? _xvartype[i] : var_t::CONTINUOUS), at(_xlower, i), at(_xupper, i),
                     at(_x0, i), at(_xf, i), at(_xtol, i));
    std::fprintf(f, "\t</states>\n\t<controls ncontrols=\"%zu\" rhorizon=\"%zu\">\n", _nControls, _urhorizon);
    for (size_t i = 0; i < _nControls; ++i)
        std::fprintf(f, "\t\t<control name=\"u%zu\" vartype=\"%c\" lower=\"%.2f\" upper=\"%.2f\"/>\n", i,
                     vt(i < _uvartype.size() ? _uvartype[i] : var_t::CONTINUOUS), at(_ulower, i), at(_uupper, i));
    std::fprintf(f, "\t</controls>\n\t<exzones nzones=\"%zu\">\n", _obstacles_raw.size());
    size_t z = 0;
    for (const border_t& b : _obstacles_raw) {
        std::fprintf(f, "\t\t<border name=\"exz%zu\" ncorners=\"%zu\">\n", z++, b.size());
        for (const corner_t& c : b)
            std::fprintf(f, "\t\t\t<corner x=\"%.2f\" y=\"%.2f\" z=\"%.2f\"/>\n", c[0], c[1], c[2]);
        std::fprintf(f, "\t\t</border>\n");
    }
    std::fprintf(f, "\t</exzones>\n\t<mexzones nzones=\"%zu\">\n", _tracks.size());
    z = 0;
    for (const track_t& tr : _tracks) {
        std::fprintf(f, "\t\t<track name=\"mexz%zu\" radius=\"%.2f\" nwaypoints=\"%zu\">\n", z++, tr.radius,
                     tr.trajectory.size());
        size_t w = 0;
        for (const traj_elem_t& wp : tr.trajectory) {
            std::fprintf(f, "\t\t\t<waypoint name=\"pt%zu\" t=\"%.2f\" ndatums=\"%zu\">\n", w++, wp.first,
                         wp.second.size());
            for (double d : wp.second) std::fprintf(f, "\t\t\t\t<datum>%.2f</datum>\n", d);
            std::fprintf(f, "\t\t\t</waypoint>\n");
        }
        std::fprintf(f, "\t\t</track>\n");
    }
    std::fprintf(f, "\t</mexzones>\n</etol>\n");
    std::fclose(f);
}

// ---- CSV ----------------------------------------------------------------------------------------------------
std::string TrajectoryOptimizer::save(traj_t* traj, std::string fp) {
    if (traj->empty()) {
        std::cout << "No Data to Save!!!" << std::endl;
        return fp;
    }
    const size_t dot = fp.find('.');
    const std::string ext = dot == std::string::npos ? std::string() : fp.substr(dot);
    struct stat st;
    while (stat(fp.c_str(), &st) != -1) {  // name taken: bump the trailing integer of the stem
        std::string stem = fp.substr(0, fp.find('.'));
        const size_t digits_at = stem.find_last_not_of("0123456789") + 1;
        int idx = digits_at == stem.size() ? 0 : std::atoi(stem.substr(digits_at).c_str());
        fp = stem.substr(0, digits_at) + std::to_string(idx + 1) + ext;
    }
    std::ofstream f(fp, std::ios::out);
    const size_t width = traj->front().second.size();
    std::string row = "time";
    for (size_t i = 0; i < width; ++i) row += ",traj" + std::to_string(i);
    f << row << "\n";
    for (size_t r = 0; r < traj->size(); ++r) {
        const traj_elem_t& e = (*traj)[r];
        row = std::to_string(e.first);  // fixed six decimals, as std::to_string prints doubles
        for (double v : e.second) row += "," + std::to_string(v);
        if (r + 1 != traj->size()) row += "\n";  // no trailing newline
        f << row;
    }
    return fp;
}

// ---- container API -----------------------------------------------------------------------------------------
void TrajectoryOptimizer::resetConfigs() {
    _nStates = 0;
    _nControls = 0;
    _nSteps = 0;
    _dt = 0.;
    _xrhorizon = _urhorizon = _rhorizon = 0;
    _xvartype.clear();
    _x0.clear();
    _xf.clear();
    _xtol.clear();
    _xlower.clear();
    _xupper.clear();
    _uvartype.clear();
    _ulower.clear();
    _uupper.clear();
    _parameters.clear();
    _obstacles_raw.clear();
    _obstacles.clear();
    _tracks.clear();
}

void TrajectoryOptimizer::printConfigs() {
    using std::cout;
    using std::endl;
    auto vec = [](const char* name, const state_t& v) {
        cout << "  " << name << ":";
        for (double d : v) cout << " " << d;
        cout << endl;
    };
    cout << "ETOL configuration" << endl;
    cout << "  nsteps: " << _nSteps << "  dt: " << _dt << endl;
    cout << "  nstates: " << _nStates << "  ncontrols: " << _nControls << "  rhorizon(x,u): " << _xrhorizon << ","
         << _urhorizon << endl;
    vec("x lower", _xlower);
    vec("x upper", _xupper);
    vec("x initial", _x0);
    vec("x terminal", _xf);
    vec("x tolerance", _xtol);
    vec("u lower", _ulower);
    vec("u upper", _uupper);
    cout << "  exclusion zones: " << _obstacles_raw.size() << endl;
    for (const border_t& b : _obstacles_raw) {
        cout << "   ";
        for (const corner_t& c : b) cout << " (" << c[0] << "," << c[1] << "," << c[2] << ")";
        cout << endl;
    }
    cout << "  moving exclusion zones: " << _tracks.size() << endl;
    for (const track_t& tr : _tracks) {
        cout << "    radius " << tr.radius << ":";
        for (const traj_elem_t& wp : tr.trajectory) {
            cout << " t=" << wp.first << " [";
            for (double d : wp.second) cout << " " << d;
            cout << " ]";
        }
        cout << endl;
    }
}

void TrajectoryOptimizer::addParams(std::list<param_t> params) {
    for (const param_t& p : params) _parameters.insert(p);  // std::map: first insertion of a name wins
}
// ---- convex partition of an exclusion zone (reference :84-159 genRegion, :161-207 calcSlopes) ------------------
namespace {
struct P2 {
    double x, y;
};
double cross(const P2& o, const P2& a, const P2& b) { return (a.x - o.x) * (b.y - o.y) - (a.y - o.y) * (b.x - o.x); }
bool inside_tri(const P2& p, const P2& a, const P2& b, const P2& c) {
    return cross(a, b, p) >= 0.0 && cross(b, c, p) >= 0.0 && cross(c, a, p) >= 0.0;
}
// counter-clockwise simple polygon -> triangles (vertex indices), ear clipping
std::vector<std::array<int, 3>> ear_clip(const std::vector<P2>& v) {
    std::vector<int> idx(v.size());
    for (size_t i = 0; i < v.size(); ++i) idx[i] = static_cast<int>(i);
    std::vector<std::array<int, 3>> tris;
    size_t guard = 0;
    while (idx.size() > 3 && guard++ < 4 * v.size() * v.size()) {
        bool clipped = false;
        for (size_t i = 0; i < idx.size(); ++i) {
            const int a = idx[(i + idx.size() - 1) % idx.size()], b = idx[i], c = idx[(i + 1) % idx.size()];
            if (cross(v[a], v[b], v[c]) <= 0.0) continue;  // reflex or degenerate corner
            bool ear = true;
            for (int q : idx)
                if (q != a && q != b && q != c && inside_tri(v[q], v[a], v[b], v[c])) {
                    ear = false;
                    break;
                }
            if (!ear) continue;
            tris.push_back({a, b, c});
            idx.erase(idx.begin() + static_cast<long>(i));
            clipped = true;
            break;
        }
        if (!clipped) {  // collinear leftovers: drop a degenerate corner and go on
            bool dropped = false;
            for (size_t i = 0; i < idx.size() && !dropped; ++i) {
                const int a = idx[(i + idx.size() - 1) % idx.size()], b = idx[i], c = idx[(i + 1) % idx.size()];
                if (cross(v[a], v[b], v[c]) == 0.0) {
                    idx.erase(idx.begin() + static_cast<long>(i));
                    dropped = true;
                }
            }
            if (!dropped) break;
        }
    }
    if (idx.size() == 3 && cross(v[idx[0]], v[idx[1]], v[idx[2]]) > 0.0) tris.push_back({idx[0], idx[1], idx[2]});
    return tris;
}
bool convex_ccw(const std::vector<int>& poly, const std::vector<P2>& v) {
    const size_t n = poly.size();
    for (size_t i = 0; i < n; ++i)
        if (cross(v[poly[i]], v[poly[(i + 1) % n]], v[poly[(i + 2) % n]]) < 0.0) return false;
    return true;
}
// Hertel-Mehlhorn: remove a diagonal whenever the union of the two pieces it separates is still convex
std::vector<std::vector<int>> merge_convex(const std::vector<std::array<int, 3>>& tris, const std::vector<P2>& v) {
    std::vector<std::vector<int>> polys;
    for (const auto& t : tris) polys.push_back({t[0], t[1], t[2]});
    bool merged = true;
    while (merged) {
        merged = false;
        for (size_t a = 0; a < polys.size() && !merged; ++a)
            for (size_t b = a + 1; b < polys.size() && !merged; ++b) {
                const std::vector<int>&A = polys[a], &B = polys[b];
                for (size_t i = 0; i < A.size() && !merged; ++i) {
                    const int p = A[i], q = A[(i + 1) % A.size()];  // edge p -> q of A; B must hold q -> p
                    for (size_t j = 0; j < B.size() && !merged; ++j) {
                        if (B[j] != q || B[(j + 1) % B.size()] != p) continue;
                        std::vector<int> u;  // A from q round to p, then B from p round to q (both without the shared edge)
                        for (size_t s = 0; s < A.size(); ++s) u.push_back(A[(i + 1 + s) % A.size()]);
                        for (size_t s = 2; s < B.size(); ++s) u.push_back(B[(j + s) % B.size()]);
                        if (!convex_ccw(u, v)) continue;
                        polys[a] = u;
                        polys.erase(polys.begin() + static_cast<long>(b));
                        merged = true;
                    }
                }
            }
    }
    return polys;
}
}  // namespace

region_t TrajectoryOptimizer::genRegion(border_t* border) {
    region_t region;
    if (border == nullptr || border->size() < 3) return region;
    std::vector<P2> v;
    for (const corner_t& c : *border) {
        if (!v.empty() && v.back().x == c.at(0) && v.back().y == c.at(1)) continue;  // repeated corner
        v.push_back({c.at(0), c.at(1)});
    }
    if (v.size() > 1 && v.front().x == v.back().x && v.front().y == v.back().y) v.pop_back();  // explicitly closed
    if (v.size() < 3) return region;
    double area2 = 0.0;
    for (size_t i = 0; i < v.size(); ++i) area2 += v[i].x * v[(i + 1) % v.size()].y - v[(i + 1) % v.size()].x * v[i].y;
    if (area2 == 0.0) return region;
    if (area2 < 0.0) std::reverse(v.begin(), v.end());  // work counter-clockwise
    for (const std::vector<int>& poly : merge_convex(ear_clip(v), v)) {
        // lower chain: counter-clockwise from the leftmost to the rightmost vertex; upper chain: the other way round,
        // stored left to right as well (the reference sorts both segments from left to right, :81-83)
        size_t il = 0, ir = 0;
        for (size_t i = 1; i < poly.size(); ++i) {
            const P2 &p = v[poly[i]], &l = v[poly[il]], &r = v[poly[ir]];
            if (p.x < l.x || (p.x == l.x && p.y < l.y)) il = i;
            if (p.x > r.x || (p.x == r.x && p.y > r.y)) ir = i;
        }
        boundary_t bd;
        for (size_t i = il;; i = (i + 1) % poly.size()) {
            bd.lower.push_back({v[poly[i]].x, v[poly[i]].y, 0.});
            if (i == ir) break;
        }
        for (size_t i = ir;; i = (i + 1) % poly.size()) {
            bd.upper.push_front({v[poly[i]].x, v[poly[i]].y, 0.});
            if (i == il) break;
        }
        region.push_back(bd);
    }
    return region;
}

void TrajectoryOptimizer::calcSlopes(const region_t& region, std::vector<seg_t>* lowers, std::vector<seg_t>* uppers) {
    lowers->clear();
    uppers->clear();
    auto chain = [](const border_t& b) {
        seg_t seg;
        for (auto it = b.begin(); it != b.end() && std::next(it) != b.end(); ++it) {
            const double dely = std::next(it)->at(1) - it->at(1), delx = std::next(it)->at(0) - it->at(0);
            edge_prop_t prop;
            prop.slope = delx == 0. ? DBL_MAX : dely / delx;  // vertical edge: DBL_MAX as in the reference (:173-176)
            prop.length = std::sqrt(delx * delx + dely * dely);
            seg.push_back(edge_t(*it, prop));
        }
        return seg;
    };
    for (const boundary_t& bd : region) {
        lowers->push_back(chain(bd.lower));
        uppers->push_back(chain(bd.upper));
    }
}

void TrajectoryOptimizer::addExclZone(border_t* border) {
    _obstacles_raw.push_back(*border);
    region_t obstacle = genRegion(border);
    if (!obstacle.empty()) _obstacles.push_back(obstacle);
}
void TrajectoryOptimizer::addAdjTrack(track_t* track) { _tracks.push_back(*track); }

void TrajectoryOptimizer::errorHandler() {
    if (_eAny != nullptr) {
        std::fprintf(stderr, "%s", _eAny->what());
        exit(EXIT_FAILURE);
    }
}

const double TrajectoryOptimizer::getScore() const { return _score; }
void TrajectoryOptimizer::setScore(const double score) { _score = score; }
state_t& TrajectoryOptimizer::getX0() { return _x0; }
void TrajectoryOptimizer::setX0(const state_t& x0) { _x0 = x0; }
state_t& TrajectoryOptimizer::getXf() { return _xf; }
void TrajectoryOptimizer::setXf(const state_t& xf) { _xf = xf; }
const size_t TrajectoryOptimizer::getNControls() const { return _nControls; }
const size_t TrajectoryOptimizer::getNStates() const { return _nStates; }
state_t& TrajectoryOptimizer::getXlower() { return _xlower; }
void TrajectoryOptimizer::setXlower(const state_t& v) { _xlower = v; }
state_t& TrajectoryOptimizer::getXupper() { return _xupper; }
void TrajectoryOptimizer::setXupper(const state_t& v) { _xupper = v; }
state_var_t& TrajectoryOptimizer::getXvartype() { return _xvartype; }
void TrajectoryOptimizer::setXvartype(const state_var_t& v) { _xvartype = v; }
const double TrajectoryOptimizer::getDt() const { return _dt; }
void TrajectoryOptimizer::setDt(double dt) { _dt = dt; }
const size_t TrajectoryOptimizer::getNSteps() const { return _nSteps; }
void TrajectoryOptimizer::setNSteps(const size_t n) { _nSteps = n; }
state_t& TrajectoryOptimizer::getXtol() { return _xtol; }
void TrajectoryOptimizer::setXtol(const state_t& v) { _xtol = v; }
state_t& TrajectoryOptimizer::getUlower() { return _ulower; }
void TrajectoryOptimizer::setUlower(const state_t& v) { _ulower = v; }
state_t& TrajectoryOptimizer::getUupper() { return _uupper; }
void TrajectoryOptimizer::setUupper(const state_t& v) { _uupper = v; }
state_var_t& TrajectoryOptimizer::getUvartype() { return _uvartype; }
void TrajectoryOptimizer::setUvartype(const state_var_t& v) { _uvartype = v; }
const size_t TrajectoryOptimizer::getUrhorizon() const { return _urhorizon; }
void TrajectoryOptimizer::setUrhorizon(const size_t n) { _urhorizon = n; }
const size_t TrajectoryOptimizer::getXrhorizon() const { return _xrhorizon; }
void TrajectoryOptimizer::setXrhorizon(const size_t n) { _xrhorizon = n; }
size_t TrajectoryOptimizer::getRhorizon() const { return std::max(_xrhorizon, _urhorizon); }
void TrajectoryOptimizer::setNControls(const size_t n) { _nControls = n; }
void TrajectoryOptimizer::setNStates(const size_t n) { _nStates = n; }
void TrajectoryOptimizer::setEqConstraints(std::vector<f_t*> c) { _eq = c; }
void TrajectoryOptimizer::setLessEqConstraints(std::vector<f_t*> c) { _lesseq = c; }
void TrajectoryOptimizer::setConstraints(std::vector<f_t*> c) { _constraints = c; }
void TrajectoryOptimizer::setGradient(std::vector<f_t*> g) { _gradient = g; }
void TrajectoryOptimizer::setObjective(f_t* objective) { _objective = objective; }
traj_t* TrajectoryOptimizer::getUtraj() { return &_utraj; }
traj_t* TrajectoryOptimizer::getXtraj() { return &_xtraj; }
const f_t* TrajectoryOptimizer::getObjective() const { return _objective; }
std::vector<f_t*>* TrajectoryOptimizer::getGradient() { return &_gradient; }
std::vector<f_t*>* TrajectoryOptimizer::getEqConstraints() { return &_eq; }
std::vector<f_t*>* TrajectoryOptimizer::getLessEqConstraints() { return &_lesseq; }
std::vector<f_t*>* TrajectoryOptimizer::getConstraints() { return &_constraints; }
std::vector<border_t>* TrajectoryOptimizer::getObstacles_Raw() { return &_obstacles_raw; }
std::list<region_t>* TrajectoryOptimizer::getObstacles() { return &_obstacles; }
std::list<track_t>* TrajectoryOptimizer::getTracks() { return &_tracks; }
paramset_t* TrajectoryOptimizer::getParams() { return &_parameters; }
bool TrajectoryOptimizer::isMaximized() const { return _maximize; }
void TrajectoryOptimizer::setMaximize(const bool m) { _maximize = m; }
size_t TrajectoryOptimizer::getNExclZones() { return _obstacles_raw.size(); }
size_t TrajectoryOptimizer::getNTracks() { return _tracks.size(); }

}  // namespace ETOL
