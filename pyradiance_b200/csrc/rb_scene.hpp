// rb_scene.hpp -- host-side scene database for the B200 ray-tracing hot path.
//
// Replaces, for this path only, the reference's scene loader:
//   .oct reader            src/radiance/common/readoct.c:35-141,195-218
//   frozen scene decoder   src/radiance/common/sceneio.c:20-109
//   portable int/float     src/radiance/common/portio.c:93-152
//   modifier resolution    src/radiance/rt/initotypes.c:112-145 (findmaterial),
//                          src/radiance/common/modobject.c:59-89 (lastmod)
//   face / cone setup      src/radiance/common/face.c:35-106, cone.c:44-218
// The output is a set of flat arrays ready to be copied to HBM (rb_device.hpp).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include <unordered_map>

namespace rb {

// Object types we know by name (index into the file's own type table is
// remapped to these at load; everything else is OT_OTHER and is rejected
// lazily, only if a ray ever needs it).
enum OType : int {
    OT_OTHER = 0,
    // surfaces
    OT_POLYGON, OT_CONE, OT_SPHERE, OT_RING, OT_CYLINDER, OT_CUP, OT_BUBBLE,
    OT_TUBE, OT_SOURCE, OT_INSTANCE, OT_MESH,
    // modifiers
    OT_ALIAS, OT_PLASTIC, OT_METAL, OT_GLASS, OT_TRANS, OT_GLOW, OT_LIGHT,
    OT_ILLUM, OT_SPOTLIGHT,
    // known but unsupported material / pattern families (for messages + flags)
    OT_DIELECTRIC, OT_INTERFACE, OT_MIST, OT_ABSDF, OT_TRANS2, OT_ANTIMATTER,
    OT_OTHER_MATERIAL, OT_PATTERN, OT_TEXTURE, OT_MIXTURE,
    OT_PLASTIC2, OT_METAL2,      // anisotropic (rt/aniso.c), with OT_TRANS2 above
    OT_BSDF,                     // BSDF (rt/m_bsdf.c), with OT_ABSDF above
    OT_NTYPES
};

bool ot_is_surface(int t);
bool ot_is_volume(int t);
bool ot_is_material(int t);
bool ot_is_light(int t);
bool ot_is_modifier(int t);
int  ot_from_name(const std::string& s);

struct Object {
    int omod = -1;              // modifier object index, -1 = void
    int otype = OT_OTHER;
    std::string tname;          // type name as in the file
    std::string name;
    std::vector<std::string> sargs;
    std::vector<double> fargs;
    bool expanded = false;      // instance / mesh replaced by world-space surfaces at load
    bool volume_obj = false;    // hits report the enclosing instance / mesh object (modifier override)
    std::vector<double> vnorm;  // mesh triangle with vertex normals: 3 world-space (unnormalised) normals
};

struct Scene {
    std::vector<std::string> header;   // info header lines (without FORMAT=)
    double cuorg[3] = {0, 0, 0};
    double cusize = 0;
    bool frozen = true;
    std::vector<std::string> srcfiles;
    std::vector<Object> objs;
    // octree: children words, 8 per node.  word >= 0: node index;
    // -1: empty; <= -2: leaf set at leafpool[-(w)-2] = count, ids ascending.
    int root = -1;
    std::vector<int> nodes;
    std::vector<int> leafpool;
    int maxdepth = 0;
    std::string error;
    std::string basedir;        // directory of the octree file: where auxiliary files (BSDF XML) are looked for last

    bool load_octree(const std::string& path);
    // last modifier named `name` defined before object `before` (-1: any)
    int lastmod(int before, const std::string& name) const;
    // findmaterial(): returns object index of the actual material or -1
    int findmaterial(int obj) const;
    std::unordered_map<std::string, int> modtab;   // name -> last modifier idx
    void index_modifiers();
    bool read_rad_text(const std::string& path);
    // replace instances (and meshes) by transformed copies of their surfaces and
    // re-build the octree over the flat surface list
    bool expand_volumes(const std::string& basedir, int depth = 0);
    int nexpanded = 0;
};

// ---- flattened (device-ready) tables -------------------------------------

// per-object header, 16 bytes: x = type | flags<<8 | nv<<16 ; y = omod ;
// z = material slot (-1 none) ; w = offset into geom[] (doubles)
enum : int { PF_TRANSP = 1, PF_HASMAT = 2 };
// x bits above the flags: 10-11 projection axis, 12 exact rectangle, 13 smooth triangle (vertex normals after the
// vertices), 14 the surface's modifier is named "Phong" (rt/rtotypes.h:13 usesPhongSmoothing)
// 15 the reference reports the enclosing instance / mesh object for a hit on this surface (o_instance.c:41-43,
// o_mesh.c:183-190: modifier override), which is not isflat() whatever the surface inside is
enum : int { PX_SMOOTH = 1 << 13, PX_PHONG = 1 << 14, PX_NOTFLAT = 1 << 15 };
enum : int {            // device primitive kinds (x & 0xff)
    PK_NONE = 0, PK_FACE, PK_SPHERE, PK_BUBBLE, PK_CONE, PK_CUP, PK_CYL,
    PK_TUBE, PK_RING, PK_UNSUPPORTED
};

// material kinds on device
enum : int {
    MK_NONE = 0, MK_PLASTIC, MK_METAL, MK_TRANS, MK_GLASS, MK_LIGHT, MK_GLOW,
    MK_ILLUM, MK_SPOT, MK_UNSUPPORTED, MK_PLASTIC2, MK_METAL2, MK_TRANS2, MK_DIELECTRIC, MK_INTERFACE,
    MK_BSDF, MK_ABSDF
};

struct MatRec {          // 96 bytes
    int kind;
    int flags;           // bit0: chain has unsupported pattern/texture
    int obj;             // object index of the material
    int nargs;
    float a[8];          // real args (as given)
    int alt;             // illum: alternate material slot (-1 = void/none)
    int pat;             // first pattern record under the material (-1 = none)
    int pad[2];          // BSDF / aBSDF: pad[0] = index into bsdfs[], pad[1] = bits of the ninth real (a[] holds eight)
    double u[3];         // plastic2 / metal2 / trans2: orientation vector, function transform applied (aniso.c:305-313);
                         // BSDF / aBSDF: the up vector, likewise (m_bsdf.c:705-712)
    double pad2;         // BSDF: thickness, scaled by the function transform
};
static_assert(sizeof(MatRec) == 96, "MatRec must be 96 bytes");

// A pattern under a material that is built as native code: brightfunc with
// gen/skybright.cal `skybr` (gensky) or gen/perezlum.cal `skybright` (gendaylight).
struct PatRec {
    int kind;            // PAT_SKYBRIGHT / PAT_PEREZLUM
    int next;            // next pattern of the same material's chain, -1 = end
    int pad[2];
    double a[10];        // A1..A10
    double xb[9];        // backward transform of the function's coordinates, already divided by its scale:
                         // D' = D . xb (func.c:455-458)
};
enum : int { PAT_SKYBRIGHT = 1, PAT_PEREZLUM = 2 };

// ---- BSDF / aBSDF materials: Klems-matrix data of one XML file (common/bsdf_m.c), rb_bsdf.cpp builds these ----
#define RB_BSDF_MAXLATS 46
struct BsdfBasis {       // bsdf_m.h ANGLE_BASIS
    int nangles, nlat;
    double tmin[RB_BSDF_MAXLATS + 1];   // lower polar bound of each latitude in degrees; tmin[nlat] closes the last one
    int nphis[RB_BSDF_MAXLATS + 1];     // azimuth count per latitude; nphis[nlat] = 0
    int pad;
};
struct BsdfComp {        // one SDSpectralDF with its single matrix component (SDMat)
    int present, ninc, nout;
    int ib, ob;          // incident / exiting basis (index into bsdfbases[])
    unsigned mtx;        // offsets into bsdfpool[] (32-bit words): float value[o * ninc + i] = mBSDF_value(o, i)
    unsigned cdf, ctot;  // unsigned cdf[ninc][nout + 1], double ctot[ninc]: make_cdist() of every incident direction
    unsigned rcdf, rctot;// unsigned rcdf[nout][ninc + 1], double rctot[nout]: the same through reciprocity
    double minProjSA, maxHemi;
};
struct BsdfRec {
    BsdfComp c[4];       // rf, rb, tf, tb (SDData; XML front / back already swapped)
    double lamb[4];      // cieY of rLambFront, rLambBack, tLambFront, tLambBack
};
enum : int { BC_RF = 0, BC_RB, BC_TF, BC_TB };

struct SrcRec {          // distant & local sources (source.h SRCREC subset)
    double sloc[3];      // direction (distant) or position
    double ss2;          // solid angle or projected area
    double ss[3][3];     // u, v, w axes (size vectors)
    double srad;         // maximum source radius
    double prox;         // glow proximity
    float  val[3];       // emitted radiance (material RGB)
    int    so;           // source object index
    int    flags;        // SF_*
    int    mat;          // material slot
    int    pad;
    double spot_aim[3];  // spotlight: unit aim vector (srcsupp.c makespot)
    float  spot_siz;     // spotlight: solid angle of the cone
    float  spot_flen;    // spotlight: focal length (length of the aim vector as given)
};
enum : int { SF_DISTANT = 1, SF_SKIP = 2, SF_PROX = 4, SF_SPOT = 8, SF_FLAT = 16,
             SF_CIRC = 32, SF_CYL = 64, SF_FOLLOW = 128 };

struct FlatScene {
    std::vector<int32_t> objhdr;     // 4 ints per object
    std::vector<double>  geom;       // packed geometry records
    std::vector<MatRec>  mats;
    std::vector<SrcRec>  srcs;
    std::vector<PatRec>  pats;
    std::vector<BsdfRec> bsdfs;      // one per BSDF file named by a BSDF / aBSDF material
    std::vector<BsdfBasis> bsdfbases;
    std::vector<uint32_t> bsdfpool;  // matrices and cumulative tables
    std::vector<int>     nodes;      // same encoding as Scene
    std::vector<int>     leafpool;
    std::vector<int>     leaf2;      // (count,0),(id, geom offset)... pairs; nodes[] index these
    int root = -1;
    int nsurf_unsupported = 0;
    std::string unsupported_note;    // first unsupported thing, for messages
    std::vector<std::string> warnings;
};

bool flatten_scene(const Scene& sc, FlatScene& fs, std::string& err);
bool rebuild_octree(Scene& sc, int objlim, int maxres, std::string& err);
bool build_octree_file(const Scene& sc, const std::string& cmdline, const std::string& oct_path, int objlim,
                       int maxres, std::string& err);
std::string find_radiance_file(const std::string& name, const std::string& basedir);
// Klems-matrix BSDF XML file -> fs.bsdfs / bsdfbases / bsdfpool (rb_bsdf.cpp); index = its slot in fs.bsdfs
bool load_klems_bsdf(const std::string& path, FlatScene& fs, int& index, std::string& err);

}  // namespace rb
