(* energy_gpu.ml - Energy.Make with total_potential_energy / total_kinetic_energy on the GPU
   (energy.ml:57-78); same ENERGY signature, so Integrator.Make, Ics.Make and Analysis.Make
   take it unchanged.  NOT COMPILED HERE.  The kinetic energy alone costs no O(N^2) sweep
   (nbk_energy with pe = NULL), the potential energy alone no momenta. *)

open Body

module Make (B : BODY) : Energy.ENERGY with type b = B.b = struct
  include Energy.Make (B)

  let total_potential_energy bs =
    Nbk.potential_energy (Nbk.ctx ()) (Nbk.pack1 B.m bs) (Nbk.pack3 B.q bs) !Base.eps2

  let total_kinetic_energy bs =
    Nbk.kinetic_energy (Nbk.ctx ()) (Nbk.pack1 B.m bs) (Nbk.pack3 B.p bs)

  let energy bs =
    let m = Nbk.pack1 B.m bs and q = Nbk.pack3 B.q bs and p = Nbk.pack3 B.p bs in
    let ke, pe = Nbk.energy (Nbk.ctx ()) m q p !Base.eps2 in
    ke +. pe
end
