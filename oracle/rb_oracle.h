/* rb_oracle.h -- TEST INFRASTRUCTURE (not product code).
 *
 * Plain-C, double-precision, recursive CPU restatement of the reference's hot
 * path (see rb_oracle.c for the file:line map).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it, and only
 * as the checker.  Parity of this restatement is PINNED against the compiled
 * reference (oracle/_ref) and the golden vectors in tests/golden/ by
 * tests/test_oracle.py.
 */
#ifndef RB_ORACLE_H
#define RB_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_scene orc_scene;

typedef struct orc_params {
    int ambounce, ambdiv, maxdepth, backvis, directvis, do_irrad;
    double minweight, dstrsrc, specthresh, specjitter;
    double ambval[3];
    int contrib;            /* -V+ */
    uint64_t seed;
    double srcsizerat;      /* -ds */
} orc_params;

/* bin function ids (same meaning as the .cal files, see rb_oracle.c) */
enum { ORC_BIN_CONST = 0, ORC_BIN_REINHARTB = 1, ORC_BIN_REINHART = 2, ORC_BIN_KLEMS_FULL = 3, ORC_BIN_HEMI = 4,
       ORC_BIN_KLEMS_HALF = 5, ORC_BIN_KLEMS_QUARTER = 6, ORC_BIN_SHIRCHIU = 7 };

typedef struct orc_result {
    double rop[3], ron[3], rot, rod;
    int32_t robj, omod;
    double value[3];
} orc_result;

typedef struct orc_counters {
    uint64_t nrays, nodes, leafents, prims, contribs;
} orc_counters;

orc_scene* orc_load(const char* octree_path, char* err, size_t errlen);
void orc_free(orc_scene* s);
int orc_num_objects(const orc_scene* s);
const char* orc_object_name(const orc_scene* s, int i);
void orc_default_params(orc_params* p, int rcontrib);
void orc_set_params(orc_scene* s, const orc_params* p);
void orc_clear_modifiers(orc_scene* s);
/* returns first column */
int orc_add_modifier(orc_scene* s, const char* name, int fn, int mf, const double n[3], const double u[3],
                     double rhs, int nbins);
int orc_num_columns(const orc_scene* s);
/* irrad: 0 none, 1 rtrace -I, 2 rcontrib -I, 3 manager */
int orc_rtrace(orc_scene* s, const double* rays, size_t nrays, int irrad, orc_result* out);
int orc_rcontrib(orc_scene* s, const double* rays, size_t nrays, int accum, int irrad, double* out);
double orc_bin(int fn, int mf, const double n[3], const double u[3], double rhs, const double dir[3]);
void orc_get_counters(const orc_scene* s, orc_counters* c);
void orc_reset_counters(orc_scene* s);
const char* orc_last_error(const orc_scene* s);
/* mode 1: localhit() runs the DEVICE walk restated on the CPU (integer cell coordinates, top-level cell
 * table, no checked-object set) instead of the recursive restatement of the reference: same answers expected */
void orc_set_walker(orc_scene* s, int mode);
unsigned long long orc_dev_nodes(const orc_scene* s);

#ifdef __cplusplus
}
#endif
#endif
