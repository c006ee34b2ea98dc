/* nbk_stubs.c — OCaml <-> libnbk glue: the C side of the `external` declarations in nbk.ml.
 *
 * NOT COMPILED IN THIS REPOSITORY'S BUILD IMAGE: the image has no OCaml toolchain (no ocaml,
 * ocamlfind, ocamlbuild, opam or dune), so <caml/...> headers are absent.  These sources are
 * what a maintainer of farr/direct-nbody adds to the reference tree; they are mechanical
 * (unbox arguments, call include/nbk.h, raise Failure on a non-zero code) and every call
 * they make is exercised here through the same C ABI from Python/C++ (tests/).
 *
 * All arrays are Bigarray.Array1 of float64, c_layout (bigarray is already linked for every
 * target of the reference, _tags:5-11), so Caml_ba_data_val gives the double* libnbk wants
 * without copying.  libnbk never calls back into OCaml and never keeps a pointer past the
 * call, so no root registration is needed; the calls are long (an O(N^2) sweep), so the
 * runtime lock could be released around them, but the reference is single-threaded
 *  and nothing is gained.
 */
#include <caml/alloc.h>
#include <caml/bigarray.h>
#include <caml/custom.h>
#include <caml/fail.h>
#include <caml/memory.h>
#include <caml/mlvalues.h>

#include "nbk.h"

#define Ctx_val(v) (*((nbk_ctx **)Data_custom_val(v)))
#define D(v) ((double *)Caml_ba_data_val(v))

static void ctx_finalize(value v) {
  if (Ctx_val(v)) nbk_destroy(Ctx_val(v));
  Ctx_val(v) = NULL;
}

static struct custom_operations ctx_ops = {"org.nbk.ctx",       ctx_finalize,
                                           custom_compare_default, custom_hash_default,
                                           custom_serialize_default, custom_deserialize_default,
                                           custom_compare_ext_default, custom_fixed_length_default};

static void check(nbk_ctx *ctx, int rc) {
  if (rc != NBK_OK) caml_failwith(nbk_last_error(ctx)); /* reference convention: exceptions */
}

CAMLprim value nbk_ml_create(value device) {
  CAMLparam1(device);
  CAMLlocal1(v);
  nbk_ctx *ctx = NULL;
  int rc = nbk_create(&ctx, Int_val(device));
  if (rc != NBK_OK) caml_failwith(nbk_last_error(NULL));
  v = caml_alloc_custom(&ctx_ops, sizeof(nbk_ctx *), 0, 1);
  Ctx_val(v) = ctx;
  CAMLreturn(v);
}

CAMLprim value nbk_ml_create_multi(value ndev) {
  CAMLparam1(ndev);
  CAMLlocal1(v);
  nbk_ctx *ctx = NULL;
  int rc = nbk_create_multi(&ctx, Int_val(ndev), NULL);
  if (rc != NBK_OK) caml_failwith(nbk_last_error(NULL));
  v = caml_alloc_custom(&ctx_ops, sizeof(nbk_ctx *), 0, 1);
  Ctx_val(v) = ctx;
  CAMLreturn(v);
}

/* external grad_v_sum : ctx -> m -> q -> eps2:float -> range(t0,t1,s0,s1) -> gsum -> unit */
CAMLprim value nbk_ml_grad_v_sum(value ctx, value m, value q, value eps2, value rng,
                                 value gsum) {
  CAMLparam5(ctx, m, q, eps2, rng);
  CAMLxparam1(gsum);
  int n = (int)Caml_ba_array_val(m)->dim[0];
  check(Ctx_val(ctx),
        nbk_grad_v_sum(Ctx_val(ctx), n, D(m), D(q), Double_val(eps2), Int_val(Field(rng, 0)),
                       Int_val(Field(rng, 1)), Int_val(Field(rng, 2)), Int_val(Field(rng, 3)),
                       D(gsum)));
  CAMLreturn(Val_unit);
}
CAMLprim value nbk_ml_grad_v_sum_bc(value *a, int n) {
  (void)n;
  return nbk_ml_grad_v_sum(a[0], a[1], a[2], a[3], a[4], a[5]);
}

/* external grad_v_dgrad_v_sum : ctx -> m -> q -> p -> eps2 -> range -> gsum -> dgsum -> unit */
CAMLprim value nbk_ml_grad_v_dgrad_v_sum(value ctx, value m, value q, value p, value eps2,
                                         value rng, value gsum, value dgsum) {
  CAMLparam5(ctx, m, q, p, eps2);
  CAMLxparam3(rng, gsum, dgsum);
  int n = (int)Caml_ba_array_val(m)->dim[0];
  check(Ctx_val(ctx),
        nbk_grad_v_dgrad_v_sum(Ctx_val(ctx), n, D(m), D(q), D(p), Double_val(eps2),
                               Int_val(Field(rng, 0)), Int_val(Field(rng, 1)),
                               Int_val(Field(rng, 2)), Int_val(Field(rng, 3)), D(gsum),
                               D(dgsum)));
  CAMLreturn(Val_unit);
}
CAMLprim value nbk_ml_grad_v_dgrad_v_sum_bc(value *a, int n) {
  (void)n;
  return nbk_ml_grad_v_dgrad_v_sum(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]);
}

/* external grad_v_dgrad_v_sum_predicted :
     ctx -> t -> m -> q -> p -> f -> df -> nt:float -> eps2 -> range -> gsum -> dgsum -> unit */
CAMLprim value nbk_ml_predicted_bc(value *a, int n) {
  (void)n;
  int nb = (int)Caml_ba_array_val(a[2])->dim[0];
  value rng = a[9];
  check(Ctx_val(a[0]),
        nbk_grad_v_dgrad_v_sum_predicted(Ctx_val(a[0]), nb, D(a[1]), D(a[2]), D(a[3]), D(a[4]),
                                         D(a[5]), D(a[6]), Double_val(a[7]), Double_val(a[8]),
                                         Int_val(Field(rng, 0)), Int_val(Field(rng, 1)),
                                         Int_val(Field(rng, 2)), Int_val(Field(rng, 3)), D(a[10]),
                                         D(a[11])));
  return Val_unit;
}
CAMLprim value nbk_ml_predicted(value a0, value a1, value a2, value a3, value a4, value a5,
                                value a6, value a7, value a8, value a9, value a10, value a11) {
  value a[12] = {a0, a1, a2, a3, a4, a5, a6, a7, a8, a9, a10, a11};
  return nbk_ml_predicted_bc(a, 12);
}

/* external energy : ctx -> m -> q -> p -> eps2 -> float * float   (ke, pe) */
CAMLprim value nbk_ml_energy(value ctx, value m, value q, value p, value eps2) {
  CAMLparam5(ctx, m, q, p, eps2);
  CAMLlocal1(res);
  double ke = 0.0, pe = 0.0;
  int n = (int)Caml_ba_array_val(m)->dim[0];
  check(Ctx_val(ctx), nbk_energy(Ctx_val(ctx), n, D(m), D(q), D(p), Double_val(eps2), &ke, &pe));
  res = caml_alloc_tuple(2);
  Store_field(res, 0, caml_copy_double(ke));
  Store_field(res, 1, caml_copy_double(pe));
  CAMLreturn(res);
}

/* external v_sum : ctx -> m -> q -> eps2 -> range -> phi -> float (sum of phi) */
CAMLprim value nbk_ml_v_sum(value ctx, value m, value q, value eps2, value rng, value phi) {
  CAMLparam5(ctx, m, q, eps2, rng);
  CAMLxparam1(phi);
  double total = 0.0;
  int n = (int)Caml_ba_array_val(m)->dim[0];
  check(Ctx_val(ctx),
        nbk_v_sum(Ctx_val(ctx), n, D(m), D(q), Double_val(eps2), Int_val(Field(rng, 0)),
                  Int_val(Field(rng, 1)), Int_val(Field(rng, 2)), Int_val(Field(rng, 3)), D(phi),
                  &total));
  CAMLreturn(caml_copy_double(total));
}
CAMLprim value nbk_ml_v_sum_bc(value *a, int n) {
  (void)n;
  return nbk_ml_v_sum(a[0], a[1], a[2], a[3], a[4], a[5]);
}

/* external kick_sum : ctx -> m -> q -> v -> eta -> dt -> range -> dv -> tmin -> unit */
CAMLprim value nbk_ml_kick_sum_bc(value *a, int n) {
  (void)n;
  int nb = (int)Caml_ba_array_val(a[1])->dim[0];
  value rng = a[6];
  check(Ctx_val(a[0]),
        nbk_kick_sum(Ctx_val(a[0]), nb, D(a[1]), D(a[2]), D(a[3]), Double_val(a[4]),
                     Double_val(a[5]), Int_val(Field(rng, 0)), Int_val(Field(rng, 1)),
                     Int_val(Field(rng, 2)), Int_val(Field(rng, 3)), D(a[7]), D(a[8])));
  return Val_unit;
}
CAMLprim value nbk_ml_kick_sum(value a0, value a1, value a2, value a3, value a4, value a5,
                               value a6, value a7, value a8) {
  value a[9] = {a0, a1, a2, a3, a4, a5, a6, a7, a8};
  return nbk_ml_kick_sum_bc(a, 9);
}

/* external grad_v_dgrad_v_sum_tangent :
     ctx -> m -> q -> p -> k:int -> dq -> dp -> eps2 -> range -> gsum_tan -> dgsum_tan -> unit */
CAMLprim value nbk_ml_tangent_bc(value *a, int n) {
  (void)n;
  int nb = (int)Caml_ba_array_val(a[1])->dim[0];
  value rng = a[8];
  check(Ctx_val(a[0]),
        nbk_grad_v_dgrad_v_sum_tangent(Ctx_val(a[0]), nb, D(a[1]), D(a[2]), D(a[3]), Int_val(a[4]),
                                       D(a[5]), D(a[6]), Double_val(a[7]), Int_val(Field(rng, 0)),
                                       Int_val(Field(rng, 1)), Int_val(Field(rng, 2)),
                                       Int_val(Field(rng, 3)), NULL, NULL, D(a[9]), D(a[10])));
  return Val_unit;
}
CAMLprim value nbk_ml_tangent(value a0, value a1, value a2, value a3, value a4, value a5,
                              value a6, value a7, value a8, value a9, value a10) {
  value a[11] = {a0, a1, a2, a3, a4, a5, a6, a7, a8, a9, a10};
  return nbk_ml_tangent_bc(a, 11);
}

/* ---- bodies resident on the device for Omelyan_advancer.OA (nbk_oa_*) ---------------------- */
/* external oa_upload : ctx -> m -> q -> p -> unit */
CAMLprim value nbk_ml_oa_upload(value ctx, value m, value q, value p) {
  CAMLparam4(ctx, m, q, p);
  int n = (int)Caml_ba_array_val(m)->dim[0];
  check(Ctx_val(ctx), nbk_oa_upload(Ctx_val(ctx), n, D(m), D(q), D(p)));
  CAMLreturn(Val_unit);
}

/* external oa_steps : ctx -> h:float -> eps2:float -> nsteps:int -> unit */
CAMLprim value nbk_ml_oa_steps(value ctx, value h, value eps2, value nsteps) {
  CAMLparam4(ctx, h, eps2, nsteps);
  check(Ctx_val(ctx), nbk_oa_steps(Ctx_val(ctx), Double_val(h), Double_val(eps2), Int_val(nsteps)));
  CAMLreturn(Val_unit);
}

/* external oa_download : ctx -> q -> p -> unit */
CAMLprim value nbk_ml_oa_download(value ctx, value q, value p) {
  CAMLparam3(ctx, q, p);
  check(Ctx_val(ctx), nbk_oa_download(Ctx_val(ctx), D(q), D(p)));
  CAMLreturn(Val_unit);
}

/* external oa_diagnostics : ctx -> eps2:float -> el -> float * float   (ke, pe) */
CAMLprim value nbk_ml_oa_diagnostics(value ctx, value eps2, value el) {
  CAMLparam3(ctx, eps2, el);
  CAMLlocal1(res);
  double ke = 0.0, pe = 0.0;
  check(Ctx_val(ctx), nbk_oa_diagnostics(Ctx_val(ctx), Double_val(eps2), &ke, &pe, D(el)));
  res = caml_alloc_tuple(2);
  Store_field(res, 0, caml_copy_double(ke));
  Store_field(res, 1, caml_copy_double(pe));
  CAMLreturn(res);
}

/* ---- Advancer.A (advancer.ml:207-428) --------------------------------------------------------- */
/* external interact_sum : ctx -> m -> q -> na:int -> eps2:float -> s -> unit
 * bodies [imin, n) re-indexed from 0, the first na of them active (advancer.ml:213-235) */
CAMLprim value nbk_ml_interact_sum(value ctx, value m, value q, value na, value eps2, value s) {
  CAMLparam5(ctx, m, q, na, eps2);
  CAMLxparam1(s);
  int n = (int)Caml_ba_array_val(m)->dim[0];
  check(Ctx_val(ctx), nbk_interact_sum(Ctx_val(ctx), n, Int_val(na), D(m), D(q), Double_val(eps2), D(s)));
  CAMLreturn(Val_unit);
}
CAMLprim value nbk_ml_interact_sum_bc(value *a, int n) {
  (void)n;
  return nbk_ml_interact_sum(a[0], a[1], a[2], a[3], a[4], a[5]);
}

/* Records resident on the device: records = Bigarray of n*35 doubles in the field order of
 * include/nbk.h (t, m, q, p, h, hmax, t0, dVdq0-2, p0, q0, q1, q2, cs). */
/* external adv_upload : ctx -> records -> unit */
CAMLprim value nbk_ml_adv_upload(value ctx, value rec) {
  CAMLparam2(ctx, rec);
  int n = (int)(Caml_ba_array_val(rec)->dim[0] / 35);
  check(Ctx_val(ctx), nbk_adv_upload(Ctx_val(ctx), n, D(rec)));
  CAMLreturn(Val_unit);
}
/* external adv_download : ctx -> i0:int -> i1:int -> records -> unit */
CAMLprim value nbk_ml_adv_download(value ctx, value i0, value i1, value rec) {
  CAMLparam4(ctx, i0, i1, rec);
  check(Ctx_val(ctx), nbk_adv_download(Ctx_val(ctx), Int_val(i0), Int_val(i1), D(rec)));
  CAMLreturn(Val_unit);
}
/* external adv_permute : ctx -> (int32, c_layout) Array1 -> unit */
CAMLprim value nbk_ml_adv_permute(value ctx, value perm) {
  CAMLparam2(ctx, perm);
  int n = (int)Caml_ba_array_val(perm)->dim[0];
  check(Ctx_val(ctx), nbk_adv_permute(Ctx_val(ctx), n, (const int *)Caml_ba_data_val(perm)));
  CAMLreturn(Val_unit);
}
/* external adv_begin_step : ctx -> imin:int -> imax:int -> hnew:float -> unit */
CAMLprim value nbk_ml_adv_begin_step(value ctx, value imin, value imax, value hnew) {
  CAMLparam4(ctx, imin, imax, hnew);
  check(Ctx_val(ctx), nbk_adv_begin_step(Ctx_val(ctx), Int_val(imin), Int_val(imax), Double_val(hnew)));
  CAMLreturn(Val_unit);
}
/* external adv_interact : ctx -> imin:int -> imax:int -> t:float -> vC:float -> eps2:float -> unit */
CAMLprim value nbk_ml_adv_interact(value ctx, value imin, value imax, value t, value vC, value eps2) {
  CAMLparam5(ctx, imin, imax, t, vC);
  CAMLxparam1(eps2);
  check(Ctx_val(ctx), nbk_adv_interact(Ctx_val(ctx), Int_val(imin), Int_val(imax), Double_val(t),
                                       Double_val(vC), Double_val(eps2)));
  CAMLreturn(Val_unit);
}
CAMLprim value nbk_ml_adv_interact_bc(value *a, int n) {
  (void)n;
  return nbk_ml_adv_interact(a[0], a[1], a[2], a[3], a[4], a[5]);
}
/* external adv_finish : ctx -> imin:int -> imax:int -> t:float -> records -> unit
 * (records receives the (imax-imin)*35 finished doubles) */
CAMLprim value nbk_ml_adv_finish(value ctx, value imin, value imax, value t, value rec) {
  CAMLparam5(ctx, imin, imax, t, rec);
  check(Ctx_val(ctx), nbk_adv_finish(Ctx_val(ctx), Int_val(imin), Int_val(imax), Double_val(t), D(rec)));
  CAMLreturn(Val_unit);
}
/* external adv_get_t : ctx -> int -> float */
CAMLprim value nbk_ml_adv_get_t(value ctx, value i) {
  CAMLparam2(ctx, i);
  double t = 0.0;
  check(Ctx_val(ctx), nbk_adv_get_t(Ctx_val(ctx), Int_val(i), &t));
  CAMLreturn(caml_copy_double(t));
}
/* external adv_init_first : ctx -> eps2:float -> unit */
CAMLprim value nbk_ml_adv_init_first(value ctx, value eps2) {
  CAMLparam2(ctx, eps2);
  check(Ctx_val(ctx), nbk_adv_init_first(Ctx_val(ctx), Double_val(eps2)));
  CAMLreturn(Val_unit);
}
/* external adv_range_max : unit -> int */
CAMLprim value nbk_ml_adv_range_max(value unit) {
  (void)unit;
  return Val_int(nbk_adv_range_max());
}
/* external adv_advance_range : ctx -> dt:float -> sf:float -> eps2:float -> hmax -> perm -> unit
 * advance_range bs imax dt sf None on the device, imax = dim hmax = dim perm (int32) */
CAMLprim value nbk_ml_adv_advance_range(value ctx, value dt, value sf, value eps2, value hmax,
                                        value perm) {
  CAMLparam5(ctx, dt, sf, eps2, hmax);
  CAMLxparam1(perm);
  int imax = (int)Caml_ba_array_val(hmax)->dim[0];
  check(Ctx_val(ctx), nbk_adv_advance_range(Ctx_val(ctx), imax, Double_val(dt), Double_val(sf),
                                            Double_val(eps2), D(hmax), (int *)Caml_ba_data_val(perm),
                                            NULL));
  CAMLreturn(Val_unit);
}
CAMLprim value nbk_ml_adv_advance_range_bc(value *a, int n) {
  (void)n;
  return nbk_ml_adv_advance_range(a[0], a[1], a[2], a[3], a[4], a[5]);
}

/* ---- Leapfrog with the bodies resident on the device (leapfrog.ml:106-158) --------------------- */
/* external lf_upload : ctx -> dt_max -> m -> q -> v -> unit */
CAMLprim value nbk_ml_lf_upload(value ctx, value dtm, value m, value q, value v) {
  CAMLparam5(ctx, dtm, m, q, v);
  int n = (int)Caml_ba_array_val(m)->dim[0];
  check(Ctx_val(ctx), nbk_lf_upload(Ctx_val(ctx), n, D(dtm), D(m), D(q), D(v)));
  CAMLreturn(Val_unit);
}
/* external lf_download : ctx -> dt_max -> q -> v -> unit */
CAMLprim value nbk_ml_lf_download(value ctx, value dtm, value q, value v) {
  CAMLparam4(ctx, dtm, q, v);
  check(Ctx_val(ctx), nbk_lf_download(Ctx_val(ctx), D(dtm), D(q), D(v)));
  CAMLreturn(Val_unit);
}
/* external lf_advance_subsystem : ctx -> dt:float -> eta:float -> t -> dt_max -> perm -> unit
 * advance_subsystem dt eta bs iend on the device, iend = dim t = dim dt_max = dim perm (int32) */
CAMLprim value nbk_ml_lf_advance_subsystem(value ctx, value dt, value eta, value t, value dtm,
                                           value perm) {
  CAMLparam5(ctx, dt, eta, t, dtm);
  CAMLxparam1(perm);
  int iend = (int)Caml_ba_array_val(t)->dim[0];
  check(Ctx_val(ctx), nbk_lf_advance_subsystem(Ctx_val(ctx), iend, Double_val(dt), Double_val(eta),
                                               D(t), D(dtm), (int *)Caml_ba_data_val(perm), NULL));
  CAMLreturn(Val_unit);
}
CAMLprim value nbk_ml_lf_advance_subsystem_bc(value *a, int n) {
  (void)n;
  return nbk_ml_lf_advance_subsystem(a[0], a[1], a[2], a[3], a[4], a[5]);
}
