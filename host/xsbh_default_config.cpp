// xsbh_default_config.cpp -- built-in deck defaults.  Same graph vocabulary as the reference's data/config/*.msp
// (main-config.msp:145-256, config_numerical_schemes.msp:8-52, config_move_particles.msp:54-152) reduced to the nodes
// that touch the short-range force path; a user deck overrides any of these keys exactly as it would upstream
// (e.g. `compute_force: lj_compute_force`, `chunk_neighbors: { config: { chunk_size: 4 } }`).
#include "xsbh_operator.h"

namespace xsbh {

static const char* kDefaults = R"YAML(
global:
  dt: 1.0e-3 ps
  rcut_inc: 1.0 ang
  timestep: 0
  max_iteration: 0
  simulation_thermostate_screen_frequency: 10

# ---- forces
compute_force_prolog: zero_force_energy
compute_force: nop
compute_force_epilog: force_to_accel
compute_all_forces_energy: [ compute_force_prolog, compute_force, compute_force_epilog ]

# ---- velocity Verlet
verlet_first_half:
  - push_f_v_r: { dt_scale: 1.0, xform_mode: INV_XFORM }
  - push_f_v: { dt_scale: 0.5, xform_mode: IDENTITY }
verlet_second_half:
  - push_f_v: { dt_scale: 0.5, xform_mode: IDENTITY }
numerical_scheme: verlet_nve
verlet_nve:
  name: NVE_scheme
  body: [ verlet_first_half, check_and_update_particles, compute_all_forces_energy, verlet_second_half ]

# ---- neighbour lists, ghosts, particle moves
chunk_neighbors:
  config: { chunk_size: 1, build_particle_offset: true, subcell_compaction: true, free_scratch_memory: false,
            scratch_mem_per_cell: 1048576, stream_prealloc_factor: 1.05 }
chunk_neighbors_impl: chunk_neighbors
update_particle_neighbors: [ amr_grid_pairs, chunk_neighbors_impl, resize_particle_locks ]
ghost_update_all_impl: ghost_update_all_no_fv
ghost_full_update: [ ghost_comm_scheme, profile_ghost_comm_scheme, ghost_update_all_impl ]
parallel_update_particles: [ migrate_cell_particles, rebuild_amr, backup_r, ghost_full_update, grid_post_processing, update_particle_neighbors ]
init_particles: [ move_particles, extend_domain, load_balance, parallel_update_particles ]
trigger_move_particles:
  rebind: { threshold: max_displ, result: trigger_move_particles }
  body: [ particle_displ_over ]
update_particles_full_body: [ move_particles, trigger_load_balance, load_balancing_if_triggered, parallel_update_particles ]
update_particles_full:
  condition: trigger_move_particles
  body: [ update_particles_full_body ]
update_particles_fast_body: [ ghost_update_r ]
update_particles_fast:
  condition: not trigger_move_particles
  body: [ update_particles_fast_body ]
check_and_update_particles: [ trigger_move_particles, update_particles_full, update_particles_fast ]

# ---- thermodynamic state
trigger_thermostate_screen: trigger_thermo_state
trigger_thermostate_compute: nop
thermostate_compute_if_triggered:
  condition: trigger_thermo_state
  body: [ default_thermostate_compute ]
thermostate_screen_if_triggered:
  condition: trigger_thermo_state
  body: [ default_thermostate_screen ]
default_thermostate_compute: simulation_thermodynamic_state
default_thermostate_screen: print_thermodynamic_state

# ---- start-up and main loop
preinit_rcut_max: [ compute_force, nbh_dist ]
init_rcut_max: [ nbh_dist ]
hw_device_init: [ mpi_comm_world, init_cuda, update_ghost_config ]
hw_device_finalize: [ finalize_cuda ]
input_data: nop            # older decks populate the system here
setup_system: input_data
begin_iteration: [ trigger_restart, trigger_analysis, trigger_snapshot, trigger_thermostate_screen, trigger_thermostate_file, trigger_thermostate_compute ]
end_iteration: [ thermostate_compute_if_triggered, thermostate_screen_if_triggered, thermostate_file_if_triggered,
                 write_restart_if_triggered, perform_analysis_if_triggered, write_snapshot_if_triggered ]
first_iteration:
  - init_particles
  - trigger_thermo_state: { force: true }
  - compute_all_forces_energy
  - default_thermostate_compute
  - default_thermostate_screen: { print_header: true }
  - next_time_step
md_loop_stop:
  rebind: { end_at: max_iteration, result: md_loop_continue }
  body: [ sim_continue ]
md_trajectory_loop:
  loop: true
  name: md_loop
  condition: md_loop_continue
  body: [ md_loop_prolog, begin_iteration, numerical_scheme, end_iteration, md_loop_epilog, next_time_step, md_loop_stop ]
simulation_epilog: nop
simulation:
  name: exaStamp_simulation
  body:
    - hw_device_init
    - init_parameters
    - preinit_rcut_max
    - domain
    - setup_system
    - place_particles
    - init_rcut_max
    - first_iteration
    - md_loop_stop
    - md_trajectory_loop
    - simulation_epilog
    - hw_device_finalize
)YAML";

Node default_config() { return parse_yaml(kDefaults); }

}  // namespace xsbh
