// SunbGemmDesc (C ABI, include/sunb200.h) -> GemmParams (kernel-side problem description).
#pragma once
#include "common.cuh"
#include "../../include/sunb200.h"

#include <string.h>

inline GemmParams sunb_desc_to_params(const SunbGemmDesc* d) {
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.M = d->M; p.N = d->N; p.K = d->K; p.taps = d->taps; p.groups = d->groups;
    p.a_goff = d->a_goff; p.c_goff = d->c_goff;
    p.a_mode = d->a_mode; p.H = d->H; p.W = d->W; p.bw = d->bw; p.bh = d->bh;
    p.A = reinterpret_cast<const bf16*>(d->A); p.lda = d->lda;
    p.Wt = reinterpret_cast<const bf16*>(d->Wt); p.ldw = d->ldw;
    p.bias = d->bias; p.bias_mod = d->bias_mod > 0 ? d->bias_mod : 1; p.bias_ld = d->bias_ld;
    p.act = d->act;
    p.resid = reinterpret_cast<const bf16*>(d->resid); p.ldr = d->ldr;
    p.row_scale = d->row_scale; p.rows_per_img = d->rows_per_img > 0 ? d->rows_per_img : 1;
    p.out = reinterpret_cast<bf16*>(d->out); p.ldc = d->ldc;
    p.out_f32 = d->out_f32; p.ldc_f32 = d->ldc_f32;
    p.out_map = d->out_map; p.oH = d->oH; p.oW = d->oW;
    p.out2 = reinterpret_cast<bf16*>(d->out2); p.ldc2 = d->ldc2;
    p.dact_aux = reinterpret_cast<const bf16*>(d->dact_aux); p.ld_aux = d->ld_aux; p.dact = d->dact;
    return p;
}
