// rb_api.cu -- the C ABI declared in include/rb200.h: context, option parser,
// calcomp stand-in for the known bin files, modifier table, compute calls.
#include "../../include/rb200.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "rb_bins.cuh"
#include "rb_engine.cuh"
#include "rb_scene.hpp"

using namespace rb;

namespace {

struct Modifier {
    std::string name;
    DBinSpec spec;
};

struct CalContext {
    // which known files have been loaded, in order (later definitions win)
    std::vector<std::string> files;
    std::map<std::string, double> vars;
    bool has(const std::string& f) const {
        for (auto& x : files) if (x == f) return true;
        return false;
    }
    // the file that provides `rbin` right now
    std::string rbin_file() const {
        for (size_t i = files.size(); i-- > 0;)
            if (files[i] == "reinhartb.cal" || files[i] == "reinhart.cal") return files[i];
        return "";
    }
};

}  // namespace

struct rb_ctx {
    int device = 0;
    std::string error;
    std::string warnings;
    rb_params prm;
    Scene scene;
    FlatScene flat;
    std::unique_ptr<Engine> eng;
    bool cuda_ok = false;
    std::string cuda_err;
    bool loaded = false;
    CalContext cal;
    std::vector<Modifier> mods;
    int ncols = 0;
    bool bins_dirty = true;
    uint64_t seed = 0x5eed5eedULL;
    size_t qcap = 0;
    uint64_t rtrace_row_base = 0;     // global index of the first ray of the next rb_rtrace call (RNG keys)
    void* user_stream = nullptr;
    std::string tmp;
};

// Engines of closed contexts, per device: streams, events, pinned counters, small queues and small scene tables
// survive the context, so a program that opens a context per call (pyradiance.rtrace() is one process per call in
// the reference) pays for them once.  Engine::recycle() decides what is worth keeping.
static std::mutex g_pool_mutex;
// (never destroyed: at process exit the CUDA runtime may be gone before static destructors run)
static std::map<int, std::vector<std::unique_ptr<Engine>>>& g_pool = *new std::map<int, std::vector<std::unique_ptr<Engine>>>();
static const size_t kPoolPerDevice = 4;

static std::unique_ptr<Engine> pool_take(int device) {
    std::lock_guard<std::mutex> lk(g_pool_mutex);
    auto& v = g_pool[device];
    if (v.empty()) return nullptr;
    std::unique_ptr<Engine> e = std::move(v.back());
    v.pop_back();
    return e;
}
static void pool_give(int device, std::unique_ptr<Engine> e) {
    if (!e || !e->recycle()) return;           // destroyed here when not worth keeping
    std::lock_guard<std::mutex> lk(g_pool_mutex);
    auto& v = g_pool[device];
    if (v.size() < kPoolPerDevice) v.push_back(std::move(e));
}

static int fail(rb_ctx* c, const std::string& msg) {
    c->error = msg;
    return -1;
}

// ------------------------------------------------------------ defaults -----
static void defaults(rb_params& p, int program) {
    memset(&p, 0, sizeof(p));
    p.rand_samp = 1; p.shadcert = .75; p.vspretest = 512; p.directvis = 1; p.srcsizerat = .2;
    p.specjitter = 1.; p.backvis = 1; p.maxdepth = -10; p.ambres = 256;
    if (program == RB_PROGRAM_RCONTRIB) {      // rt/rcontrib.c:24-58
        p.dstrsrc = 0.9; p.shadthresh = 0.; p.directrelay = 3; p.specthresh = .02;
        p.minweight = 2e-3; p.ambacc = 0.; p.ambdiv = 350; p.ambssamp = 0; p.ambounce = 1;
    } else {                                   // rt/raycalls.c:121-155
        p.dstrsrc = 0.0; p.shadthresh = .03; p.directrelay = 2; p.specthresh = .15;
        p.minweight = 1e-4; p.ambacc = 0.1; p.ambdiv = 1024; p.ambssamp = 512; p.ambounce = 0;
    }
}

// ---------------------------------------------------- tiny expressions -----
// Enough of calcomp (common/calexpr.c) for -bn / eval(): numbers, + - * / ^,
// parentheses, variables set with -e/-p, and the bin-count constants of the
// known files.
namespace {
struct ExprParser {
    const char* s;
    rb_ctx* c;
    std::string err;
    void ws() { while (isspace((unsigned char)*s)) s++; }
    bool lookup(const std::string& id, double& v) {
        auto it = c->cal.vars.find(id);
        if (id == "PI") { v = RB_PI; return true; }
        auto mf = [&]() { auto m = c->cal.vars.find("MF"); return m == c->cal.vars.end() ? 1 : (int)m->second; };
        if (id == "Nrbins") {
            std::string f = c->cal.rbin_file();
            if (f == "reinhartb.cal") { v = rb_reinhart_nbins(mf()); return true; }
            if (f == "reinhart.cal") {
                if (c->cal.vars.find("MF") == c->cal.vars.end()) { err = "reinhart.cal needs MF to be set (-e MF:n)"; return false; }
                v = rb_reinhart_nbins(mf()) + 1; return true;
            }
            err = "Nrbins: no Reinhart bin file loaded"; return false;
        }
        if (id == "Nkbins" && c->cal.has("klems_full.cal")) { v = 145; return true; }
        if (id == "Nkhbins" && c->cal.has("klems_half.cal")) { v = 77; return true; }
        if (id == "Nkqbins" && c->cal.has("klems_quarter.cal")) { v = 41; return true; }
        if (it != c->cal.vars.end()) { v = it->second; return true; }
        if (id == "MF" && c->cal.has("reinhartb.cal")) { v = 1; return true; }
        if (id == "RHS") { v = 1; return true; }
        err = "undefined variable '" + id + "' (only the known bin-function files are understood)";
        return false;
    }
    bool primary(double& v) {
        ws();
        if (*s == '(') { s++; if (!expr(v)) return false; ws(); if (*s != ')') { err = "missing )"; return false; } s++; return true; }
        if (*s == '-') { s++; if (!primary(v)) return false; v = -v; return true; }
        if (*s == '+') { s++; return primary(v); }
        if (isdigit((unsigned char)*s) || *s == '.') { char* e; v = strtod(s, &e); if (e == s) { err = "bad number"; return false; } s = e; return true; }
        if (isalpha((unsigned char)*s) || *s == '_') {
            std::string id;
            while (isalnum((unsigned char)*s) || *s == '_' || *s == '.' || *s == '`') id.push_back(*s++);
            ws();
            if (*s == '(') {
                std::vector<double> args; s++;
                for (;;) { double a; if (!expr(a)) return false; args.push_back(a); ws(); if (*s == ',') { s++; continue; } break; }
                if (*s != ')') { err = "missing )"; return false; }
                s++;
                if (id == "floor" && args.size() == 1) { v = floor(args[0]); return true; }
                if (id == "sqrt" && args.size() == 1) { v = sqrt(args[0]); return true; }
                if (id == "if" && args.size() == 3) { v = args[0] > 0 ? args[1] : args[2]; return true; }
                err = "unsupported function '" + id + "'"; return false;
            }
            return lookup(id, v);
        }
        err = std::string("syntax error at '") + s + "'";
        return false;
    }
    bool power(double& v) {
        if (!primary(v)) return false;
        ws();
        if (*s == '^') { s++; double e; if (!power(e)) return false; v = pow(v, e); }
        return true;
    }
    bool term(double& v) {
        if (!power(v)) return false;
        for (;;) {
            ws();
            if (*s == '*') { s++; double r; if (!power(r)) return false; v *= r; }
            else if (*s == '/') { s++; double r; if (!power(r)) return false; v /= r; }
            else return true;
        }
    }
    bool expr(double& v) {
        if (!term(v)) return false;
        for (;;) {
            ws();
            if (*s == '+') { s++; double r; if (!term(r)) return false; v += r; }
            else if (*s == '-') { s++; double r; if (!term(r)) return false; v -= r; }
            else return true;
        }
    }
};
}  // namespace

static bool eval_expr(rb_ctx* c, const std::string& e, double& v, std::string& err) {
    ExprParser p{e.c_str(), c, ""};
    if (!p.expr(v)) { err = p.err; return false; }
    p.ws();
    if (*p.s) { err = std::string("syntax error at '") + p.s + "'"; return false; }
    return true;
}

// "MF:4", "MF=4,rNx=0", "MF=4;Ux=1" ...  (rt/func.c:76-119 set_eparams, scompile)
static bool apply_assignments(rb_ctx* c, const std::string& s, std::map<std::string, double>& vars, std::string& err) {
    size_t i = 0;
    while (i < s.size()) {
        while (i < s.size() && (isspace((unsigned char)s[i]) || s[i] == ',' || s[i] == ';')) i++;
        if (i >= s.size()) break;
        size_t j = i;
        while (j < s.size() && s[j] != '=' && s[j] != ':' && s[j] != ',' && s[j] != ';') j++;
        if (j >= s.size() || (s[j] != '=' && s[j] != ':')) { err = "bad parameter assignment '" + s.substr(i) + "'"; return false; }
        std::string name = s.substr(i, j - i);
        while (!name.empty() && isspace((unsigned char)name.back())) name.pop_back();
        size_t k = j + 1;
        int depth = 0;
        size_t e = k;
        while (e < s.size() && (depth > 0 || (s[e] != ',' && s[e] != ';'))) {
            if (s[e] == '(') depth++; else if (s[e] == ')') depth--;
            e++;
        }
        double v;
        std::map<std::string, double> saved = c->cal.vars;
        for (auto& kv : vars) c->cal.vars[kv.first] = kv.second;
        bool ok = eval_expr(c, s.substr(k, e - k), v, err);
        c->cal.vars = saved;
        if (!ok) { err = "parameter '" + name + "': " + err; return false; }
        vars[name] = v;
        i = e;
    }
    return true;
}

static std::string strip(const std::string& s) {
    std::string o;
    for (char ch : s) if (!isspace((unsigned char)ch)) o.push_back(ch);
    return o;
}

static bool parse_numlist(rb_ctx* c, const std::string& s, std::vector<double>& out, std::string& err) {
    size_t i = 0;
    while (i <= s.size()) {
        size_t j = i; int depth = 0;
        while (j < s.size() && (depth > 0 || s[j] != ',')) { if (s[j] == '(') depth++; else if (s[j] == ')') depth--; j++; }
        double v;
        if (!eval_expr(c, s.substr(i, j - i), v, err)) return false;
        out.push_back(v);
        i = j + 1;
        if (j >= s.size()) break;
    }
    return true;
}

// Recognise the bin expression (see rb_bins.cuh) -- anything else is rejected.
static bool make_binspec(rb_ctx* c, const std::string& params, const std::string& binexpr_in, int nbins,
                         DBinSpec& b, std::string& err) {
    memset(&b, 0, sizeof(b));
    std::map<std::string, double> vars = c->cal.vars;
    if (!apply_assignments(c, params, vars, err)) return false;
    auto var = [&](const char* n, double dflt) { auto it = vars.find(n); return it == vars.end() ? dflt : it->second; };
    std::string e = strip(binexpr_in.empty() ? "0" : binexpr_in);
    b.nbins = nbins; b.rhs = var("RHS", 1); b.mf = 1;
    // constant?
    {
        char* ep; double v = strtod(e.c_str(), &ep);
        if (ep != e.c_str() && *ep == 0) {
            int bc = (int)(v + 1.5);
            if (bc != 1) { err = "illegal non-zero constant for bin (" + binexpr_in + ")"; return false; }
            b.fn = BIN_CONST; b.cbin = 0; b.nbins = 1;
            return true;
        }
    }
    if (nbins <= 0) { err = "unspecified or illegal bin count"; return false; }
    if (e == "rbin") {
        std::string f = c->cal.rbin_file();
        if (f == "reinhartb.cal") {
            b.fn = BIN_REINHARTB; b.mf = (int)var("MF", 1);
            b.n[0] = var("rNx", 0); b.n[1] = var("rNy", 0); b.n[2] = var("rNz", -1);
            b.u[0] = var("Ux", 0); b.u[1] = var("Uy", 1); b.u[2] = var("Uz", 0);
        } else if (f == "reinhart.cal") {
            if (vars.find("MF") == vars.end()) { err = "reinhart.cal needs MF to be set (-e MF:n)"; return false; }
            b.fn = BIN_REINHART; b.mf = (int)var("MF", 1);
        } else { err = "bin expression 'rbin' needs -f reinhartb.cal or -f reinhart.cal"; return false; }
        if (b.mf < 1) { err = "illegal MF"; return false; }
        return true;
    }
    if (e == "scbin") {               // Shirley-Chiu square bins (util/disk2square.cal), rfluxmtx h=scN
        if (!c->cal.has("disk2square.cal")) { err = "bin expression 'scbin' needs -f disk2square.cal"; return false; }
        b.fn = BIN_SHIRCHIU; b.mf = (int)var("SCdim", 0);
        if (b.mf < 1) { err = "scbin needs SCdim to be set (-p SCdim=n,...)"; return false; }
        b.n[0] = var("rNx", 0); b.n[1] = var("rNy", 0); b.n[2] = var("rNz", -1);
        b.u[0] = var("Ux", 0); b.u[1] = var("Uy", 1); b.u[2] = var("Uz", 0);
        return true;
    }
    struct KForm { const char* pre; const char* file; int fn; };
    const KForm kf[3] = {{"kbin", "klems_full.cal", BIN_KLEMS_FULL}, {"khbin", "klems_half.cal", BIN_KLEMS_HALF},
                         {"kqbin", "klems_quarter.cal", BIN_KLEMS_QUARTER}};
    for (const auto& k : kf) {
        std::string pre = k.pre;
        if (e.compare(0, pre.size(), pre) != 0) continue;
        std::string rest = e.substr(pre.size());
        if (!c->cal.has(k.file)) { err = "bin expression '" + binexpr_in + "' needs -f " + k.file; return false; }
        std::vector<double> a;
        if (rest == "N") a = {0, -1, 0, 0, 0, 1};
        else if (rest == "E") a = {-1, 0, 0, 0, 0, 1};
        else if (rest == "S") a = {0, 1, 0, 0, 0, 1};
        else if (rest == "W") a = {1, 0, 0, 0, 0, 1};
        else if (rest == "D") a = {0, 0, -1, 0, 1, 0};
        else if (rest.size() > 2 && rest[0] == '(' && rest.back() == ')') {
            std::map<std::string, double> saved = c->cal.vars;
            c->cal.vars = vars;
            bool ok = parse_numlist(c, rest.substr(1, rest.size() - 2), a, err);
            c->cal.vars = saved;
            if (!ok) return false;
        }
        if (a.size() != 6) continue;
        b.fn = k.fn;
        for (int i = 0; i < 3; i++) { b.n[i] = a[i]; b.u[i] = a[3 + i]; }
        return true;
    }
    // uniform hemisphere: if(-Dx*a-Dy*b-Dz*c,0,-1)   (util/rfluxmtx.c:465-472)
    if (e.compare(0, 3, "if(") == 0 && e.size() > 9 && e.compare(e.size() - 6, 6, ",0,-1)") == 0) {
        std::string m = e.substr(3, e.size() - 9);
        double nv[3]; bool ok = true;
        const char* names[3] = {"-Dx*", "-Dy*", "-Dz*"};
        size_t pos = 0;
        for (int i = 0; i < 3 && ok; i++) {
            if (m.compare(pos, 4, names[i]) != 0) { ok = false; break; }
            pos += 4;
            size_t nxt = (i < 2) ? m.find(names[i + 1], pos) : m.size();
            if (nxt == std::string::npos) { ok = false; break; }
            std::string num = m.substr(pos, nxt - pos);
            std::map<std::string, double> saved = c->cal.vars;
            c->cal.vars = vars;
            std::string e2;
            ok = eval_expr(c, num, nv[i], e2);
            c->cal.vars = saved;
            pos = nxt;
        }
        if (ok) {
            b.fn = BIN_HEMI; b.n[0] = nv[0]; b.n[1] = nv[1]; b.n[2] = nv[2];
            return true;
        }
    }
    err = "unsupported bin expression '" + binexpr_in +
          "': only rbin (reinhartb.cal / reinhart.cal), kbin/khbin/kqbin (klems_*.cal), scbin (disk2square.cal), "
          "if(-Dx*nx-Dy*ny-Dz*nz,0,-1) and constant 0 are built as native code (no .cal interpreter)";
    return false;
}

static bool rebuild_bins(rb_ctx* c, std::string& err) {
    std::vector<DBinSpec> specs;
    std::map<std::string, int> slot;
    for (size_t i = 0; i < c->mods.size(); i++) { specs.push_back(c->mods[i].spec); slot[c->mods[i].name] = (int)i; }
    std::vector<int> otrack(c->scene.objs.size(), -1);
    // tracked name = name of the object's immediate modifier (rt/rcontrib.c:287)
    for (size_t i = 0; i < c->scene.objs.size(); i++) {
        int om = c->scene.objs[i].omod;
        if (om < 0) continue;
        auto it = slot.find(c->scene.objs[om].name);
        if (it != slot.end()) otrack[i] = it->second;
    }
    if (!c->eng->set_bins(specs, otrack, c->ncols, err)) return false;
    c->bins_dirty = false;
    return true;
}

static DParams device_params(const rb_ctx* c, bool contrib, bool need_values) {
    DParams P;
    memset(&P, 0, sizeof(P));
    const rb_params& p = c->prm;
    P.ambounce = p.ambounce; P.ambdiv = p.ambdiv; P.maxdepth = p.maxdepth; P.backvis = p.backvis;
    P.directvis = p.directvis; P.do_irrad = p.do_irrad; P.contrib = contrib; P.need_values = need_values;
    P.minweight = (float)p.minweight;
    for (int k = 0; k < 3; k++) P.ambval[k] = (float)p.ambval[k];
    P.dstrsrc = p.dstrsrc; P.specthresh = p.specthresh; P.specjitter = p.specjitter; P.srcsizerat = p.srcsizerat;
    P.seed = c->seed;
    return P;
}

// Options that would change the reference's result and are not restated fail BY NAME (no silent
// change of meaning, no CPU fallback).  `p` = the parameters in effect for this call (rcontrib's
// forced -dt 0 -as 0 -aa 0 of rt/rcmain.c:164-171 already applied).
static int check_params(rb_ctx* c, const rb_params& p) {
    if (p.ambounce > 0 && p.ambacc > 1e-6)
        return fail(c, "unsupported option: the irradiance cache (-aa > 0) is not built; use -aa 0 with -ab > 0");
    if (p.cextinction[0] > 0 || p.cextinction[1] > 0 || p.cextinction[2] > 0)
        return fail(c, "unsupported option: participating media (-me) are not built");
    if (p.maxdepth <= 0 && p.minweight <= 0)
        return fail(c, "zero ray weight in Russian roulette");
    if (p.specjitter > 1.5)
        return fail(c, "unsupported option: -ss > 1.5 (several specular samples per hit, normal.c:374-386) is not built");
    // rlvl and rdepth travel in 6 bits each (pack_info, rb_shade.cuh)
    if (p.maxdepth > 63 || p.maxdepth < -63)
        return fail(c, "unsupported option: -lr beyond +-63 (the ray level is carried in 6 bits)");
    if (p.ambounce > 63)
        return fail(c, "unsupported option: -ab above 63");
    // ambsupersamp() (rt/ambcomp.c:323-346) runs when ns*ns > MINADIV^2 = 49 and ambssamp*wt + .5 >= 4*ns
    // (ambcomp.c:413-420); both sides shrink with the ray weight, the left one faster, so a setting
    // that does not trigger at wt = 1 never triggers.  It needs the values of the first pass, which a
    // forward wavefront does not have: reject instead of silently skipping it.
    if (p.ambounce > 0 && p.ambssamp > 0 && p.ambdiv > 0) {
        const int ns = std::max(1, (int)(sqrt((double)p.ambdiv) + .5));
        if (ns * ns > 49 && (int)(p.ambssamp + .5) >= 4 * ns)
            return fail(c, "unsupported option: ambient super-sampling (-as " + std::to_string(p.ambssamp) +
                           " with -ad " + std::to_string(p.ambdiv) + ") is not built; pass -as 0");
    }
    // direct()'s adaptive shadow testing (rt/source.c:461-556): sources are sorted by potential and
    // tested until the untested tail falls below -dt times the value accumulated so far, the rest is
    // added by running hit statistics -- order dependent and tied to the recursion.  The first
    // MINSHADCNT = 2 candidates are always tested (source.c:490), so with at most two candidates
    // per shading point the threshold never acts and every setting gives the -dt 0 result.
    if (p.shadthresh > 0 && c->loaded) {
        int ncand = 0;
        for (const SrcRec& sr : c->flat.srcs) {
            if (sr.flags & SF_SKIP) continue;
            // a local source splits into partitions when -ds > 0 (srcsamp.c:36-144)
            ncand += ((sr.flags & SF_DISTANT) || p.srcsizerat <= 1e-6) ? 1 : 3;
        }
        if (ncand > 2)
            return fail(c, "unsupported option: -dt " + std::to_string(p.shadthresh) +
                           " with more than two light-source candidates per point (adaptive shadow testing, "
                           "source.c:461-556, is not built: every source is tested); pass -dt 0");
    }
    return 0;
}

// ------------------------------------------------------------- C ABI -------
extern "C" {

const char* rb_version(void) { return "pyradiance_b200 0.1 (sm_100a)"; }
int rb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

rb_ctx* rb_create(int cuda_device) {
    rb_ctx* c = new rb_ctx();
    c->device = cuda_device;
    defaults(c->prm, RB_PROGRAM_RTRACE);
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        c->cuda_ok = false;
        c->cuda_err = std::string("no CUDA device available (") + cudaGetErrorString(e) +
                      "); this library has no CPU fallback";
        cudaGetLastError();
    } else if (cuda_device < 0 || cuda_device >= n) {
        c->cuda_ok = false;
        c->cuda_err = "CUDA device " + std::to_string(cuda_device) + " out of range";
    } else {
        c->cuda_ok = true;
        c->eng = pool_take(cuda_device);
        if (!c->eng) c->eng.reset(new Engine(cuda_device));
    }
    return c;
}

void rb_destroy(rb_ctx* c) {
    if (!c) return;
    if (c->eng) pool_give(c->device, std::move(c->eng));
    delete c;
}

const char* rb_last_error(rb_ctx* c) { return c ? c->error.c_str() : "null context"; }

int rb_set_defaults(rb_ctx* c, int program) { defaults(c->prm, program); return 0; }
int rb_get_params(rb_ctx* c, rb_params* out) { *out = c->prm; return 0; }
int rb_set_params(rb_ctx* c, const rb_params* in) { c->prm = *in; return 0; }
int rb_set_seed(rb_ctx* c, uint64_t seed) { c->seed = seed; return 0; }
int rb_set_queue_capacity(rb_ctx* c, size_t n) { c->qcap = n; if (c->eng) c->eng->set_queue_capacity(n); return 0; }

static bool isflt(const char* s) { char* e; strtod(s, &e); return e != s && *e == 0; }
static bool isint(const char* s) { char* e; strtol(s, &e, 10); return e != s && *e == 0; }

int rb_set_option(rb_ctx* c, int ac, const char* const* av) {
    rb_params& p = c->prm;
    if (ac < 1 || !av[0] || av[0][0] != '-') return -1;
    auto checkf = [&](int ol, int n) { if (av[0][ol]) return false; if (ac - 1 < n) return false; for (int i = 1; i <= n; i++) if (!isflt(av[i])) return false; return true; };
    auto checki = [&](int ol) { return !av[0][ol] && ac > 1 && isint(av[1]); };
    auto cbool = [&](int ol, int& var) { switch (av[0][ol]) { case 0: var = !var; return true; case '+': case '1': var = 1; return true; case '-': case '0': var = 0; return true; } return false; };
    switch (av[0][1]) {
    case 'u': return cbool(2, p.rand_samp) ? 0 : -1;
    case 'b': if (av[0][2] == 'v') return cbool(3, p.backvis) ? 0 : -1; break;
    case 'd':
        switch (av[0][2]) {
        case 't': if (!checkf(3, 1)) return -1; p.shadthresh = atof(av[1]); return 1;
        case 'c': if (!checkf(3, 1)) return -1; p.shadcert = atof(av[1]); return 1;
        case 'j': if (!checkf(3, 1)) return -1; p.dstrsrc = atof(av[1]); return 1;
        case 'r': if (!checki(3)) return -1; p.directrelay = atoi(av[1]); return 1;
        case 'p': if (!checki(3)) return -1; p.vspretest = atoi(av[1]); return 1;
        case 'v': return cbool(3, p.directvis) ? 0 : -1;
        case 's': if (!checkf(3, 1)) return -1; p.srcsizerat = atof(av[1]); return 1;
        }
        break;
    case 's':
        switch (av[0][2]) {
        case 't': if (!checkf(3, 1)) return -1; p.specthresh = atof(av[1]); return 1;
        case 's': if (!checkf(3, 1)) return -1; p.specjitter = atof(av[1]); return 1;
        }
        break;
    case 'l':
        switch (av[0][2]) {
        case 'r': if (!checki(3)) return -1; p.maxdepth = atoi(av[1]); return 1;
        case 'w': if (!checkf(3, 1)) return -1; p.minweight = atof(av[1]); return 1;
        }
        break;
    case 'i': return cbool(2, p.do_irrad) ? 0 : -1;
    case 'a':
        switch (av[0][2]) {
        case 'v': if (!checkf(3, 3)) return -1; p.ambval[0] = atof(av[1]); p.ambval[1] = atof(av[2]); p.ambval[2] = atof(av[3]); return 3;
        case 'w': if (!checki(3)) return -1; p.ambvwt = atoi(av[1]); return 1;
        case 'a': if (!checkf(3, 1)) return -1; p.ambacc = atof(av[1]); return 1;
        case 'r': if (!checki(3)) return -1; p.ambres = atoi(av[1]); return 1;
        case 'd': if (!checki(3)) return -1; p.ambdiv = atoi(av[1]); return 1;
        case 's': if (!checki(3)) return -1; p.ambssamp = atoi(av[1]); return 1;
        case 'b': if (!checki(3)) return -1; p.ambounce = atoi(av[1]); return 1;
        }
        break;
    case 'm':
        switch (av[0][2]) {
        case 'e': if (!checkf(3, 3)) return -1; for (int k = 0; k < 3; k++) p.cextinction[k] = atof(av[1 + k]); return 3;
        case 'a': if (!checkf(3, 3)) return -1; for (int k = 0; k < 3; k++) p.salbedo[k] = atof(av[1 + k]); return 3;
        case 'g': if (!checkf(3, 1)) return -1; p.seccg = atof(av[1]); return 1;
        case 's': if (!checkf(3, 1)) return -1; p.ssampdist = atof(av[1]); return 1;
        }
        break;
    case 'f':
        if (av[0][2] || ac < 2) return -1;
        return rb_cal_load(c, av[1]) == 0 ? 1 : -2;
    case 'e':
        if (av[0][2] || ac < 2) return -1;
        if (!strchr(av[1], '=') && !strchr(av[1], ':')) break;
        return rb_cal_set(c, av[1]) == 0 ? 1 : -2;
    }
    return -1;
}

int rb_load_octree(rb_ctx* c, const char* path) {
    c->loaded = false;
    if (!c->scene.load_octree(path)) return fail(c, c->scene.error);
    std::string err;
    if (!flatten_scene(c->scene, c->flat, err)) return fail(c, err);
    c->warnings.clear();
    for (auto& w : c->flat.warnings) c->warnings += w + "\n";
    // the scene is parsed (queries work) but nothing can be traced without a GPU
    if (!c->cuda_ok) return fail(c, c->cuda_err);
    if (!c->eng->upload_scene(c->flat, c->scene, err)) return fail(c, err);
    c->loaded = true;
    c->bins_dirty = true;
    return 0;
}

int rb_save_octree(rb_ctx* c, const char* path) {
    if (c->scene.objs.empty()) return fail(c, "no octree loaded");
    std::string err;
    if (!rb::build_octree_file(c->scene, "rb_save_octree", path, 6, 16384, err)) return fail(c, err);
    return 0;
}

int rb_num_objects(rb_ctx* c) { return (int)c->scene.objs.size(); }
const char* rb_object_name(rb_ctx* c, int i) { return (i >= 0 && i < (int)c->scene.objs.size()) ? c->scene.objs[i].name.c_str() : ""; }
const char* rb_object_type(rb_ctx* c, int i) { return (i >= 0 && i < (int)c->scene.objs.size()) ? c->scene.objs[i].tname.c_str() : ""; }
int rb_object_modifier(rb_ctx* c, int i) { return (i >= 0 && i < (int)c->scene.objs.size()) ? c->scene.objs[i].omod : -1; }
int rb_num_header_lines(rb_ctx* c) { return (int)c->scene.header.size(); }
const char* rb_header_line(rb_ctx* c, int i) { return (i >= 0 && i < (int)c->scene.header.size()) ? c->scene.header[i].c_str() : ""; }
const char* rb_scene_warnings(rb_ctx* c) { return c->warnings.c_str(); }

int rb_cal_load(rb_ctx* c, const char* calfile) {
    std::string f = calfile;
    size_t sl = f.rfind('/');
    if (sl != std::string::npos) f = f.substr(sl + 1);
    static const char* known[] = {"reinhartb.cal", "reinhart.cal", "klems_full.cal", "klems_half.cal",
                                  "klems_quarter.cal", "disk2square.cal", "rayinit.cal"};
    for (const char* k : known)
        if (f == k) { c->cal.files.push_back(f); return 0; }
    return fail(c, "unsupported function file \"" + std::string(calfile) +
                       "\": only reinhartb.cal, reinhart.cal, klems_{full,half,quarter}.cal and disk2square.cal are built as native bin functions (no .cal interpreter, no CPU fallback)");
}

int rb_cal_set(rb_ctx* c, const char* assignments) {
    std::string err;
    if (!apply_assignments(c, assignments, c->cal.vars, err)) return fail(c, err);
    return 0;
}

int rb_cal_eval(rb_ctx* c, const char* expr, double* value) {
    std::string err;
    if (!eval_expr(c, expr, *value, err)) return fail(c, std::string("cannot evaluate '") + expr + "': " + err);
    return 0;
}

int rb_clear_modifiers(rb_ctx* c) { c->mods.clear(); c->ncols = 0; c->bins_dirty = true; return 0; }

int rb_add_modifier(rb_ctx* c, const char* modname, const char* params, const char* binexpr, int nbins) {
    std::string name = modname ? modname : "";
    if (name.empty() || name == "void") return fail(c, "cannot track 'void' modifier");
    for (auto& m : c->mods) if (m.name == name) return fail(c, "duplicate modifier '" + name + "'");
    Modifier m; m.name = name;
    std::string err;
    if (!make_binspec(c, params ? params : "", binexpr ? binexpr : "0", nbins, m.spec, err))
        return fail(c, "modifier '" + name + "': " + err);
    m.spec.col0 = c->ncols;
    c->ncols += m.spec.nbins;
    c->mods.push_back(m);
    c->bins_dirty = true;
    return m.spec.col0;
}

int rb_num_columns(rb_ctx* c) { return c->ncols; }

int rb_bin_of_direction(rb_ctx* c, int mi, const double dir[3], double* binval) {
    if (mi < 0 || mi >= (int)c->mods.size()) return fail(c, "bad modifier index");
    *binval = rb_eval_bin(c->mods[mi].spec, dir);
    return 0;
}

int rb_rcontrib(rb_ctx* c, const double* rays, size_t nrays, int accum, unsigned flags, uint64_t row_base,
                void* out, size_t out_floats) {
    if (!c->cuda_ok) return fail(c, c->cuda_err);
    if (!c->loaded) return fail(c, "no octree loaded");
    if (c->mods.empty()) return fail(c, "missing required modifier argument");
    if (accum <= 0) return fail(c, "unsupported option: -c 0 (single accumulated record) is not built yet");
    // rcontrib overrides (rt/rcmain.c:164-171, rxcmain.cpp:154-156): -dt 0 -as 0 -aa 0, whatever the caller set
    rb_params eff = c->prm;
    eff.shadthresh = 0; eff.ambssamp = 0; eff.ambacc = 0;
    if (check_params(c, eff) < 0) return -1;
    std::string err;
    if (c->bins_dirty && !rebuild_bins(c, err)) return fail(c, err);
    size_t nrec = (nrays + accum - 1) / (size_t)accum;
    if (out_floats < nrec * (size_t)c->ncols * 3) return fail(c, "output buffer too small");
    TraceJob job;
    job.rays = rays; job.nrays = nrays; job.accum = accum;
    job.rays_on_device = flags & RB_FLAG_RAYS_ON_DEVICE;
    job.cmat = out; job.cmat_on_device = flags & RB_FLAG_OUT_ON_DEVICE;
    job.cmat_double = flags & RB_FLAG_OUT_DOUBLE;
    job.irrad = flags & RB_FLAG_IRRAD_MASK; job.lim_dist = flags & RB_FLAG_LIMDIST;
    job.row_base = row_base;
    rb_params saved = c->prm;
    c->prm = eff;
    DParams P = device_params(c, flags & RB_FLAG_CONTRIB, flags & RB_FLAG_CONTRIB);
    c->prm = saved;
    if (!c->eng->run(job, P, err)) return fail(c, err);
    return 0;
}

int rb_rtrace(rb_ctx* c, const double* rays, size_t nrays, unsigned flags, double* values, rb_ray_result* results) {
    if (!c->cuda_ok) return fail(c, c->cuda_err);
    if (!c->loaded) return fail(c, "no octree loaded");
    if (check_params(c, c->prm) < 0) return -1;
    static_assert(sizeof(rb_ray_result) == sizeof(RayResult), "result layout");
    TraceJob job;
    job.rays = rays; job.nrays = nrays; job.accum = 1;
    job.rays_on_device = flags & RB_FLAG_RAYS_ON_DEVICE;
    job.values = values; job.results = (RayResult*)results;
    job.irrad = flags & RB_FLAG_IRRAD_MASK; job.lim_dist = flags & RB_FLAG_LIMDIST;
    job.row_base = c->rtrace_row_base;
    DParams P = device_params(c, false, values != nullptr);
    std::string err;
    if (!c->eng->run(job, P, err)) return fail(c, err);
    return 0;
}

int rb_set_row_base(rb_ctx* c, uint64_t first_ray) { c->rtrace_row_base = first_ray; return 0; }

int rb_get_stats(rb_ctx* c, rb_stats* o) {
    memset(o, 0, sizeof(*o));
    if (!c->eng) return 0;
    const EngineStats& s = c->eng->stats;
    o->nrays = s.nrays; o->nodes = s.nodes; o->leafents = s.leafents; o->prims = s.prims; o->contribs = s.contribs;
    o->launches = s.launches; o->wave_launches = s.wave_launches; o->waves = s.waves; o->batches = s.batches;
    o->retries = s.retries; o->badbin = s.badbin; o->kernel_ms = s.kernel_ms; o->wave_ms = s.wave_ms; o->shade_ms = s.shade_ms;
    return 0;
}
int rb_reset_stats(rb_ctx* c) { if (c->eng) c->eng->stats = EngineStats(); return 0; }

int rb_set_stream(rb_ctx* c, void* s) {
    if (!c->eng) return fail(c, c->cuda_err);
    c->eng->set_stream((cudaStream_t)s);
    c->user_stream = s;
    return 0;
}

void* rb_device_alloc(rb_ctx* c, size_t bytes) {
    if (!c->cuda_ok) { c->error = c->cuda_err; return nullptr; }
    cudaSetDevice(c->device);
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) { c->error = cudaGetErrorString(e); return nullptr; }
    return p;
}
int rb_device_free(rb_ctx* c, void* p) { cudaSetDevice(c->device); return cudaFree(p) == cudaSuccess ? 0 : fail(c, "cudaFree failed"); }
int rb_device_upload(rb_ctx* c, void* dst, const void* src, size_t bytes) {
    cudaSetDevice(c->device);
    cudaError_t e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
    return e == cudaSuccess ? 0 : fail(c, cudaGetErrorString(e));
}
int rb_device_download(rb_ctx* c, void* dst, const void* src, size_t bytes) {
    cudaSetDevice(c->device);
    cudaError_t e = cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost);
    return e == cudaSuccess ? 0 : fail(c, cudaGetErrorString(e));
}
int rb_mtx_multiply(rb_ctx* c, const float* a, size_t nrows, size_t ninner, const float* b, size_t ncols, float* out,
                    unsigned flags, double* kernel_ms) {
    if (!c->cuda_ok) return fail(c, c->cuda_err);
    std::string err;
    cudaStream_t st = c->user_stream ? (cudaStream_t)c->user_stream : (cudaStream_t)0;
    if (!rb::mtx_multiply(c->device, st, a, nrows, ninner, b, ncols, out, (flags & RB_MTX_A_ON_DEVICE) != 0,
                          (flags & RB_MTX_B_ON_DEVICE) != 0, (flags & RB_MTX_OUT_ON_DEVICE) != 0, kernel_ms, err))
        return fail(c, err);
    return 0;
}

int rb_view_rays(rb_ctx* c, const rb_view* view, int xres, int yres, int repeat, double pj, uint64_t seed, double* out,
                 unsigned flags) {
    if (!c->cuda_ok) return fail(c, c->cuda_err);
    if (!view || xres <= 0 || yres <= 0 || repeat <= 0 || !out) return fail(c, "rb_view_rays: bad arguments");
    if (!strchr("vlchas", view->type)) return fail(c, "unknown view type");
    std::string err;
    if (!rb::view_rays(c->device, (cudaStream_t)c->user_stream, *view, xres, yres, repeat, pj, seed, out,
                       (flags & RB_FLAG_OUT_ON_DEVICE) != 0, err))
        return fail(c, err);
    return 0;
}

int rb_ipc_export(rb_ctx* c, void* dptr, void* handle_out) {
    static_assert(sizeof(cudaIpcMemHandle_t) == RB_IPC_HANDLE_BYTES, "handle size");
    if (!c->cuda_ok) return fail(c, c->cuda_err);
    cudaSetDevice(c->device);
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, dptr);
    if (e != cudaSuccess) return fail(c, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
    memcpy(handle_out, &h, sizeof(h));
    return 0;
}
int rb_ipc_open(rb_ctx* c, const void* handle, void** dptr_out) {
    if (!c->cuda_ok) return fail(c, c->cuda_err);
    cudaSetDevice(c->device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(c, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e)); }
    *dptr_out = p;
    return 0;
}
int rb_ipc_close(rb_ctx* c, void* dptr) {
    if (!c->cuda_ok) return fail(c, c->cuda_err);
    cudaSetDevice(c->device);
    cudaError_t e = cudaIpcCloseMemHandle(dptr);
    return e == cudaSuccess ? 0 : fail(c, std::string("cudaIpcCloseMemHandle: ") + cudaGetErrorString(e));
}

int rb_host_register(rb_ctx* c, void* p, size_t bytes) {
    if (!c->cuda_ok) return fail(c, c->cuda_err);
    cudaSetDevice(c->device);
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterDefault);
    return e == cudaSuccess ? 0 : fail(c, cudaGetErrorString(e));
}
int rb_host_unregister(rb_ctx* c, void* p) {
    if (!c->cuda_ok) return fail(c, c->cuda_err);
    cudaSetDevice(c->device);
    cudaError_t e = cudaHostUnregister(p);
    return e == cudaSuccess ? 0 : fail(c, cudaGetErrorString(e));
}
int rb_device_sync(rb_ctx* c) {
    cudaSetDevice(c->device);
    cudaError_t e = cudaDeviceSynchronize();
    return e == cudaSuccess ? 0 : fail(c, cudaGetErrorString(e));
}

}  // extern "C"
