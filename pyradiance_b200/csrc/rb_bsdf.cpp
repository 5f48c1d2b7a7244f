// rb_bsdf.cpp -- host side of the BSDF / aBSDF materials: a Klems-matrix BSDF XML file -> device tables.
//
// Restated reference functions (grayscale matrix data only):
//   SDloadFile                 src/radiance/common/bsdf.c:166-243  (what is accepted, insignificant components dropped)
//   SDloadMtx                  src/radiance/common/bsdf_m.c:662-726
//   load_angle_basis           src/radiance/common/bsdf_m.c:316-368 (+ the three built-in Klems bases :31-63)
//   load_bsdf_data/get_extrema src/radiance/common/bsdf_m.c:370-533
//   extract_diffuse/subtract_min/mBSDF_color   :293-314,596-660
//   make_cdist                 src/radiance/common/bsdf_m.c:802-831 -- built here for EVERY incident direction (and for
//                              every exiting direction, the reciprocity case) instead of lazily with a cache list
// The reference reads the file with ezxml; the element tree below is a small reader of its own (elements, text,
// comments / declarations skipped, attributes ignored -- the loader never looks at one).
// Tensor-tree files and colour (CIE-X / CIE-Z) blocks are refused by name.
#include "rb_scene.hpp"

#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>

namespace rb {
namespace {

struct XNode {
    std::string name, text;
    std::vector<std::unique_ptr<XNode>> kids;
    const XNode* child(const char* n) const {
        for (const auto& k : kids) if (k->name == n) return k.get();
        return nullptr;
    }
    std::vector<const XNode*> children(const char* n) const {
        std::vector<const XNode*> v;
        for (const auto& k : kids) if (k->name == n) v.push_back(k.get());
        return v;
    }
};
static const XNode XEMPTY;
static const XNode& sub(const XNode* p, const char* n) { const XNode* c = p ? p->child(n) : nullptr; return c ? *c : XEMPTY; }
static std::string trimmed(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && isspace((unsigned char)s[a])) a++;
    while (b > a && isspace((unsigned char)s[b - 1])) b--;
    return s.substr(a, b - a);
}
static bool ieq(const std::string& a, const char* b) { return strcasecmp(a.c_str(), b) == 0; }

static bool parse_xml(const std::string& s, XNode& root, std::string& err) {
    std::vector<XNode*> st;
    size_t i = s.find('<');
    bool have_root = false;
    while (i != std::string::npos && i < s.size()) {
        if (s[i] != '<') {                       // character data up to the next tag
            size_t j = s.find('<', i);
            if (j == std::string::npos) j = s.size();
            if (!st.empty()) st.back()->text.append(s, i, j - i);
            i = j;
            continue;
        }
        if (s.compare(i, 4, "<!--") == 0) { size_t j = s.find("-->", i + 4); if (j == std::string::npos) break; i = j + 3; continue; }
        if (s.compare(i, 9, "<![CDATA[") == 0) {
            size_t j = s.find("]]>", i + 9);
            if (j == std::string::npos) { err = "unclosed <![CDATA["; return false; }
            if (!st.empty()) st.back()->text.append(s, i + 9, j - i - 9);
            i = j + 3; continue;
        }
        if (s.compare(i, 2, "<?") == 0) { size_t j = s.find("?>", i + 2); if (j == std::string::npos) break; i = j + 2; continue; }
        if (s.compare(i, 2, "<!") == 0) { size_t j = s.find('>', i + 2); if (j == std::string::npos) break; i = j + 1; continue; }
        size_t j = s.find('>', i);
        if (j == std::string::npos) { err = "missing >"; return false; }
        if (s[i + 1] == '/') {                   // closing tag
            if (st.empty()) { err = "unexpected closing tag"; return false; }
            st.pop_back();
            i = j + 1;
            if (st.empty() && have_root) break;
            continue;
        }
        size_t k = i + 1;
        while (k < j && !isspace((unsigned char)s[k]) && s[k] != '/') k++;
        std::string name = s.substr(i + 1, k - i - 1);
        size_t colon = name.find(':');           // a namespace prefix is not part of the names the loader asks for
        if (colon != std::string::npos) name = name.substr(colon + 1);
        const bool selfclose = s[j - 1] == '/';
        XNode* nd;
        if (!have_root) { root.name = name; nd = &root; have_root = true; }
        else if (st.empty()) break;
        else { st.back()->kids.emplace_back(new XNode()); nd = st.back()->kids.back().get(); nd->name = name; }
        if (!selfclose) st.push_back(nd);
        i = j + 1;
    }
    if (!have_root) { err = "root tag missing"; return false; }
    if (!st.empty()) { err = "unclosed tag <" + st.back()->name + ">"; return false; }
    return true;
}

struct HostBasis { std::string name; std::vector<double> tmin; std::vector<int> nphis; int nangles = 0; };

static double basis_ohm(const HostBasis& b, int ndx) {          // io_getohm
    int li = 0;
    for (; ndx >= b.nphis[li]; li++) ndx -= b.nphis[li];
    const double D2R = M_PI / 180.;
    const double c0 = cos(D2R * b.tmin[li]), c1 = cos(D2R * b.tmin[li + 1]);
    return M_PI * (c0 * c0 - c1 * c1) / (double)b.nphis[li];
}

struct HostComp {
    bool present = false;
    int ninc = 0, nout = 0, ib = -1, ob = -1;
    std::vector<float> v;                        // mBSDF_value(o, i) = v[o * ninc + i]
    double minProjSA = 0, maxHemi = 0;
};

// mBSDF_color() for grayscale data: the value with its position-specific perturbation
static float perturbed(const HostComp& c, int i, int o) {
    float coef = c.v[(size_t)o * c.ninc + i];
    double d = 2 * c.ninc / (i + .22545) + 4 * c.nout / (o + .70281);
    d -= (int)d;
    coef *= 1. + 6e-4 * (d - .5);
    return coef;
}

}  // namespace

bool load_klems_bsdf(const std::string& path, FlatScene& fs, int& index, std::string& err) {
    std::ifstream f(path, std::ios::binary);
    if (!f) { err = "Cannot open BSDF \"" + path + "\""; return false; }
    std::stringstream ss; ss << f.rdbuf();
    XNode root; std::string xe;
    if (!parse_xml(ss.str(), root, xe)) { err = "BSDF \"" + path + "\" " + xe; return false; }
    if (root.name != "WindowElement") { err = "BSDF \"" + path + "\": top level node not 'WindowElement'"; return false; }
    if (const XNode* ft = root.child("FileType"))
        if (trimmed(ft->text) != "BSDF") { err = "XML \"" + path + "\": wrong FileType (must be 'BSDF')"; return false; }
    const XNode* wtl = sub(&root, "Optical").child("Layer");
    if (!wtl) { err = "BSDF \"" + path + "\": no optical layers"; return false; }
    const XNode& dd = sub(wtl, "DataDefinition");
    const std::string ids = trimmed(sub(&dd, "IncidentDataStructure").text);
    if (strncasecmp(ids.c_str(), "TensorTree", 10) == 0) {
        err = "BSDF \"" + path + "\": tensor-tree data is not built (Klems-matrix files only)"; return false;
    }
    if (ids.empty()) { err = "BSDF \"" + path + "\": missing IncidentDataStructure"; return false; }
    bool row_in;
    if (ieq(ids, "Rows")) row_in = true;
    else if (ieq(ids, "Columns")) row_in = false;
    else { err = "BSDF \"" + path + "\": unsupported IncidentDataStructure"; return false; }

    std::vector<HostBasis> bases = {
        {"LBNL/Klems Full", {0., 5., 15., 25., 35., 45., 55., 65., 75., 90.}, {1, 8, 16, 20, 24, 24, 24, 16, 12, 0}, 145},
        {"LBNL/Klems Half", {0., 6.5, 19.5, 32.5, 45.5, 58.5, 71.5, 90.}, {1, 8, 12, 16, 20, 12, 8, 0}, 77},
        {"LBNL/Klems Quarter", {0., 9., 27., 45., 63., 90.}, {1, 8, 12, 12, 8, 0}, 41}};
    for (const XNode* wab : dd.children("AngleBasis")) {
        const std::string name = trimmed(sub(wab, "AngleBasisName").text);
        if (name.empty()) continue;
        bool known = false;
        for (const auto& b : bases) if (ieq(b.name, name.c_str())) known = true;
        if (known) continue;                     // "assume it's the same"
        if (bases.size() >= 7) { err = "Out of angle bases reading '" + name + "'"; return false; }
        HostBasis hb; hb.name = name; hb.tmin.push_back(0.);
        int i = 0;
        for (const XNode* wbb : wab->children("AngleBasisBlock")) {
            if (i >= RB_BSDF_MAXLATS) { err = "Too many latitudes for '" + name + "'"; return false; }
            const XNode& tb = sub(wbb, "ThetaBounds");
            const double up = atof(trimmed(sub(&tb, "UpperTheta").text).c_str());
            if (i) {
                double a = atof(trimmed(sub(&tb, "LowerTheta").text).c_str());
                const double b = hb.tmin[i];
                if (b != 0) a = a / b - 1.;
                if (!((a <= 1e-6) & (a >= -1e-6))) { err = "Theta values disagree in '" + name + "'"; return false; }
            }
            hb.tmin.push_back(up);
            const int np = atoi(trimmed(sub(wbb, "nPhis").text).c_str());
            if (np <= 0 || (np == 1 && hb.tmin[i] > 1e-6)) { err = "Illegal phi count in '" + name + "'"; return false; }
            hb.nphis.push_back(np); hb.nangles += np;
            i++;
        }
        hb.nphis.push_back(0);
        bases.push_back(std::move(hb));
    }
    auto find_basis = [&](const std::string& n) { for (size_t k = bases.size(); k--;) if (ieq(bases[k].name, n.c_str())) return (int)k; return -1; };

    HostComp comp[4];                            // rf, rb, tf, tb
    for (const XNode* wld : wtl->children("WavelengthData")) {
        const std::string cnm = trimmed(sub(wld, "Wavelength").text);
        if (ieq(cnm, "CIE-X") || ieq(cnm, "CIE-Z")) {
            err = "BSDF \"" + path + "\": colour (CIE-X / CIE-Z) data is not built; use the Visible-only file"; return false;
        }
        if (!ieq(cnm, "Visible")) continue;
        for (const XNode* wdb : wld->children("WavelengthDataBlock")) {
            const std::string dir = trimmed(sub(wdb, "WavelengthDataDirection").text);
            int ci;                              // front and back are reversed from WINDOW 6 orientations
            if (ieq(dir, "Transmission Front")) ci = 3;
            else if (ieq(dir, "Transmission Back")) ci = 2;
            else if (ieq(dir, "Reflection Front")) ci = 1;
            else if (ieq(dir, "Reflection Back")) ci = 0;
            else continue;
            const std::string cb = trimmed(sub(wdb, "ColumnAngleBasis").text), rbn = trimmed(sub(wdb, "RowAngleBasis").text);
            if (cb.empty()) { err = "Missing column basis for BSDF '" + path + "'"; return false; }
            const int inbi = find_basis(cb);
            if (inbi < 0) { err = "Undefined ColumnAngleBasis '" + cb + "'"; return false; }
            if (rbn.empty()) { err = "Missing row basis for BSDF '" + path + "'"; return false; }
            const int outbi = find_basis(rbn);
            if (outbi < 0) { err = "Undefined RowAngleBasis '" + rbn + "'"; return false; }
            HostComp& c = comp[ci];
            c = HostComp();
            c.present = true; c.ib = inbi; c.ob = outbi;
            c.ninc = bases[inbi].nangles; c.nout = bases[outbi].nangles;
            c.v.assign((size_t)c.ninc * c.nout, 0.f);
            const std::string& sd = sub(wdb, "ScatteringData").text;
            const char* p = sd.c_str();
            while (isspace((unsigned char)*p)) p++;
            if (!*p) { err = "Missing BSDF ScatteringData in '" + path + "'"; return false; }
            for (int i = 0; i < c.ninc * c.nout; i++) {
                char* e = nullptr;
                double val = strtod(p, &e);
                if (e == p) { err = "Bad/missing BSDF ScatteringData in '" + path + "'"; return false; }
                p = e;
                while (isspace((unsigned char)*p)) p++;
                if (*p == ',') p++;
                if (val < 0) val = 0;            // don't allow negative values
                if (row_in) { const int r = i / c.nout, col = i - r * c.nout; c.v[(size_t)col * c.ninc + r] = (float)val; }
                else c.v[i] = (float)val;
            }
            // get_extrema()
            c.minProjSA = M_PI; c.maxHemi = 0.;
            std::vector<double> ohma(c.nout);
            for (int o = c.nout; o--;) if ((ohma[o] = basis_ohm(bases[outbi], o)) < c.minProjSA) c.minProjSA = ohma[o];
            for (int i = c.ninc; i--;) {
                double hemi = 0.;
                for (int o = c.nout; o--;) hemi += ohma[o] * c.v[(size_t)o * c.ninc + i];
                if (hemi > c.maxHemi) c.maxHemi = hemi;
            }
            if (inbi != outbi)
                for (int i = c.ninc; i--;) { const double ohm = basis_ohm(bases[inbi], i); if (ohm < c.minProjSA) c.minProjSA = ohm; }
        }
    }
    // extract_diffuse() -> subtract_min(), grayscale
    double lamb[4] = {0, 0, 0, 0};               // rLambFront, rLambBack, tLambFront, tLambBack
    auto extract = [&](HostComp& c) -> double {
        if (!c.present) return 0.;
        float ymin = 1e10f;
        for (int i = 0; i < c.ninc; i++)
            for (int o = 0; o < c.nout; o++) { const float v = perturbed(c, i, o); if (v < ymin) ymin = v; }
        if (ymin <= .01 / M_PI) return 0.;
        for (auto& v : c.v) v -= ymin;
        const double cieY = M_PI * ymin;
        c.maxHemi -= cieY;
        return cieY;
    };
    lamb[0] = extract(comp[0]);
    lamb[1] = extract(comp[1]);
    lamb[2] = extract(comp[2]);
    if (comp[3].present) {
        lamb[3] = extract(comp[3]);
        if (!comp[2].present) lamb[2] = lamb[3];
    } else if (comp[2].present)
        lamb[3] = lamb[2];
    for (auto& c : comp) if (c.present && c.maxHemi <= .001) c = HostComp();     // insignificant components

    // ---- device tables ----
    BsdfRec rec; memset(&rec, 0, sizeof(rec));
    for (int k = 0; k < 4; k++) rec.lamb[k] = lamb[k];
    std::vector<int> bmap(bases.size(), -1);
    auto dev_basis = [&](int bi) {
        if (bmap[bi] >= 0) return bmap[bi];
        BsdfBasis db; memset(&db, 0, sizeof(db));
        const HostBasis& hb = bases[bi];
        db.nangles = hb.nangles; db.nlat = (int)hb.nphis.size() - 1;
        for (size_t k = 0; k < hb.tmin.size(); k++) db.tmin[k] = hb.tmin[k];
        for (size_t k = 0; k < hb.nphis.size(); k++) db.nphis[k] = hb.nphis[k];
        bmap[bi] = (int)fs.bsdfbases.size();
        fs.bsdfbases.push_back(db);
        return bmap[bi];
    };
    auto& pool = fs.bsdfpool;
    auto put_doubles = [&](const std::vector<double>& v) {
        if (pool.size() & 1) pool.push_back(0);
        const unsigned off = (unsigned)pool.size();
        pool.resize(pool.size() + 2 * v.size());
        if (!v.empty()) memcpy(&pool[off], v.data(), v.size() * sizeof(double));
        return off;
    };
    for (int k = 0; k < 4; k++) {
        const HostComp& c = comp[k];
        BsdfComp& d = rec.c[k];
        if (!c.present) continue;
        d.present = 1; d.ninc = c.ninc; d.nout = c.nout;
        d.ib = dev_basis(c.ib); d.ob = dev_basis(c.ob);
        d.minProjSA = c.minProjSA; d.maxHemi = c.maxHemi;
        d.mtx = (unsigned)pool.size();
        pool.resize(pool.size() + c.v.size());
        memcpy(&pool[d.mtx], c.v.data(), c.v.size() * sizeof(float));
        // make_cdist() for every incident direction (forward) and every exiting one (reverse = reciprocity)
        for (int rev = 0; rev < 2; rev++) {
            const int nidx = rev ? c.nout : c.ninc, calen = rev ? c.ninc : c.nout;
            std::vector<double> ohm(calen), tot(nidx);
            for (int o = 0; o < calen; o++) ohm[o] = basis_ohm(bases[rev ? c.ib : c.ob], o);
            const unsigned off = (unsigned)pool.size();
            pool.resize(pool.size() + (size_t)nidx * (calen + 1));
            std::vector<double> cm(calen + 1);
            for (int ix = 0; ix < nidx; ix++) {
                cm[0] = 0.;
                for (int o = 0; o < calen; o++) {
                    const float val = rev ? c.v[(size_t)ix * c.ninc + o] : c.v[(size_t)o * c.ninc + ix];
                    cm[o + 1] = val * ohm[o];
                    cm[o + 1] += cm[o];
                }
                tot[ix] = cm[calen];
                const double scale = (double)0xffffffffu / cm[calen];
                uint32_t* carr = &pool[off + (size_t)ix * (calen + 1)];
                carr[0] = 0;
                for (int o = 1; o < calen; o++) {
                    const double x = scale * cm[o] + .5;         // cTotal = 0 never gets sampled (SDsampComponent)
                    carr[o] = (x >= 4294967295.0) ? 0xffffffffu : (x > 0 ? (uint32_t)x : 0u);
                }
                carr[calen] = 0xffffffffu;
            }
            const unsigned toff = put_doubles(tot);
            if (rev) { d.rcdf = off; d.rctot = toff; } else { d.cdf = off; d.ctot = toff; }
        }
    }
    index = (int)fs.bsdfs.size();
    fs.bsdfs.push_back(rec);
    return true;
}

}  // namespace rb
