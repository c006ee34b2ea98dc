(* energy_gpu.ml — Energy.Make with total_potential_energy / total_kinetic_energy on the GPU
   (energy.ml:57-78); same ENERGY signature, so Integrator.Make, Ics.Make and Analysis.Make
   take it unchanged.  NOT COMPILED HERE. *)

open Body

module Make (B : BODY) : Energy.ENERGY with type b = B.b = struct
  include Energy.Make (B)

  let arrays bs = (Nbk.pack1 B.m bs, Nbk.pack3 B.q bs, Nbk.pack3 B.p bs)

  let total_potential_energy bs =
    let m, q, p = arrays bs in
    snd (Nbk.energy (Nbk.ctx ()) m q p !Base.eps2)

  let total_kinetic_energy bs =
    let m, q, p = arrays bs in
    fst (Nbk.energy (Nbk.ctx ()) m q p !Base.eps2)

  let energy bs =
    let m, q, p = arrays bs in
    let ke, pe = Nbk.energy (Nbk.ctx ()) m q p !Base.eps2 in
    ke +. pe
end
