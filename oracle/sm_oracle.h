/* oracle/sm_oracle.h -- C API of the CPU restatement of the reference's hot path.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load libsmoracle.so; the product never does.
 *
 * Pinning: the reference ships no tests or golden vectors for this path, so the
 * restatement is pinned against the reference's own code run here (oracle/_ref/libsmref.so, the
 * reference headers compiled verbatim): tests/test_oracle_port.py drives both with identical inputs
 * and requires bit-identical columns, particle states and counters, and tests/golden/ holds
 * vectors generated from oracle/_ref by tests/golden/make_golden.py.
 */
#ifndef SM_ORACLE_H
#define SM_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct smo_soil {   /* numeric SurfParam fields, surface.h:11-39 */
  int32_t transports, erodes, cascades, abrades;
  float density, porosity, solubility, equrate, friction, erosionrate, maxdiff, settling,
      suspension, abrasion;
} smo_soil;

typedef struct smo_stats {
  int64_t steps, sweeps, exit_oob, exit_evap, exit_stall;
  double seconds;
} smo_stats;

typedef struct smo_hydro {   /* counters of one hydrology call (flood phase or seep pass) */
  int64_t floods;        /* flood() calls that passed the volume/spill guard, nested particles included */
  int64_t nested;        /* particles spawned by the water-table cascade (water.h:243-256) */
  int64_t nested_steps;  /* their particle-steps */
  int64_t transfers;     /* partial water-table transfers (water.h:260-272) */
  int64_t cells;         /* seep pass: cells visited */
} smo_hydro;

void smo_init(int dimx, int dimy, int scale, int nsoils, const smo_soil* soils);
void smo_set_columns(const int64_t* offsets, const int32_t* type, const double* size, const double* saturation);
int64_t smo_nsections(void);
void smo_get_columns(int64_t* offsets, int32_t* type, double* size, double* floor, double* saturation);
void smo_heights(double* out);
void smo_get_frequency(float* wfreq, float* wtrack, float* windfreq);
void smo_set_frequency(const float* wfreq, const float* wtrack, const float* windfreq);
void smo_frequency_update(void);

double smo_height_i(int x, int y);
double smo_height_f(float x, float y);
int smo_surface(int x, int y);
void smo_normal(int x, int y, float* out3);
void smo_add(int x, int y, double size, int type);
double smo_remove(int x, int y, double h);
void smo_cascade(float x, float y, int transferloop);

/* lockstep: every sweep, every live particle in ascending index does move() && interact() */
void smo_water_begin(int n, const float* xy);
int smo_water_sweep(smo_stats* st);
void smo_water_state(float* pos2, float* speed2, double* volume, double* sediment, int32_t* contains, int32_t* alive);
void smo_water_run(int n, const float* xy, int max_sweeps, smo_stats* st);
/* pooling hydrology: flood() of every finished particle of the last lockstep
 * water batch in ascending index, each flood atomic (water.h:123-145); the per-frame full-grid seep pass
 * (water.h:335-343) */
void smo_water_flood(smo_hydro* out);
void smo_seep(smo_hydro* out);
void smo_wind_begin(int n, const float* xy);
int smo_wind_sweep(smo_stats* st);
void smo_wind_state(float* pos2, float* speed3, double* height, double* sediment, int32_t* contains, int32_t* alive);
void smo_wind_run(int n, const float* xy, int max_sweeps, smo_stats* st);
/* sequential: each particle runs to completion before the next is spawned (SoilMachine.cpp:288-307) */
void smo_water_seq(int n, const float* xy, smo_stats* st);
void smo_water_seq_full(int n, const float* xy, int flags, smo_stats* st, smo_hydro* out);
void smo_wind_seq(int n, const float* xy, smo_stats* st);

#ifdef __cplusplus
}
#endif
#endif
