// tests/facade_demo.cpp -- the reference's frame loop (SoilMachine.cpp:283-320, GL-free part) written
// against include/soilmachine/soilmachine.hpp.  Compiled (and linked against the product library) by
// the CPU test-suite; executed on the GPU box, where it prints the counters of two frames.
#include <cstdio>
#include <cstdlib>
#include "../include/soilmachine/soilmachine.hpp"
using namespace soilmachine;

int SIZEX = 128, SIZEY = 128, SCALE = 80, NWATER = 300, NWIND = 150, SEED = 42;
struct DummyVertexpool {} vertexpool;

int main(int argc, char** argv) {
  srand(SEED);                                               // SoilMachine.cpp:41
  if (argc > 1) {                                            // loadsoil(parse::option["soil"]), SoilMachine.cpp:43-45
    try {
      WorldEntry w = loadsoil(argv[1]);
      SCALE = w.scale;                                       // SIZEX/SIZEY stay small for the demo
    } catch (const SoilFileError& e) { printf("%s\n", e.what()); return 2; }
  } else {
  // what loadsoil("soil/rocksand.soil") leaves in the tables (io.h:7-230), abbreviated
  SurfParam rock; rock.name = "Rock"; rock.transports = rock.erodes = rock.cascades = rock.abrades = 1;
  rock.density = 0.95f; rock.solubility = 1.0f; rock.equrate = 0.1f; rock.friction = 0.15f; rock.maxdiff = 0.01f; rock.settling = 0.1f;
  SurfParam sand = rock; sand.name = "Red Sand"; sand.transports = sand.erodes = sand.cascades = sand.abrades = 2;
  sand.density = 0.4f; sand.porosity = 0.8f; sand.friction = 0.1f; sand.maxdiff = 0.005f; sand.settling = 0.05f; sand.suspension = 0.01f;
  soils.push_back(rock); soilmap["Rock"] = 1; soils.push_back(sand); soilmap["Red Sand"] = 2;
  SurfLayer l0(1); l0.bias = 0.5f; l0.scale = 0.8f; l0.octaves = 8; l0.lacunarity = 2; l0.gain = 0.5f; l0.frequency = 1;
  SurfLayer l1(2); l1.bias = 0.0f; l1.scale = 0.4f; l1.octaves = 6; l1.lacunarity = 2; l1.gain = 0.4f; l1.frequency = 2;
  layers.push_back(l0); layers.push_back(l1);
  }
  WaterParticle::init(SIZEX, SIZEY); WindParticle::init(SIZEX, SIZEY);   // :47-48 (host mirrors of the maps)
  try {
    Layermap map(SEED, ivec2(SIZEX, SIZEY), vertexpool, SCALE);   // :83
    for (int frame = 0; frame < 2; frame++) {                // the body of Tiny::loop, :287-320
      sm_stats w = WaterParticle::run(map, vertexpool, NWATER);
      sm_hydro_stats fl = WaterParticle::flood_batch(map, vertexpool);   // :296, for the whole batch
      sm_hydro_stats se = WaterParticle::seep(map, vertexpool);    // :300-301
      sm_stats d = WindParticle::run(map, vertexpool, NWIND);
      WaterParticle::mapfrequency(map);
      WaterParticle::resetfrequency(map);
      float wf = 0.f, df = 0.f;                              // the texture loops of :315-326 read the mirrors
      for (int i = 0; i < SIZEX * SIZEY; i++) { wf += WaterParticle::frequency[i]; df += WindParticle::frequency[i]; }
      printf("  mean water frequency %.6g, mean wind frequency %.6g\n", wf / (SIZEX * SIZEY), df / (SIZEX * SIZEY));
      printf("frame %d: water %lld steps in %lld sweeps, %lld floods, seep pass over %lld cells, wind %lld steps; h(5,5)=%.17g surface=%zu\n", frame,
             (long long)w.steps, (long long)w.sweeps, (long long)fl.floods, (long long)se.cells, (long long)d.steps,
             map.height(ivec2(5, 5)), map.surface(ivec2(5, 5)));
    }
    map.add(ivec2(3, 3), map.pool.get(0.01, soilmap["Red Sand"]));   // legacy per-cell idiom
    double left = map.remove(ivec2(3, 3), 0.005);
    Particle::cascade(vec2(3.2f, 3.4f), map, vertexpool, 1);
    printf("legacy ops ok, leftover %.3g\n", left);
  } catch (const Error& e) {
    printf("soilmachine error %d: %s\n", e.code, e.what());
    return e.code == SM_ERR_NOGPU ? 77 : 1;
  }
  return 0;
}
