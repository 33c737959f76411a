/* examples/demo.c - the reference's `cargo run --example demo` (examples/demo.rs:20-130) against the C ABI of include/tsgpu.h:
 * setup_params(3); an 8-cell MemoryTrace W(0,42) W(1,100) R(0) R(1) W(0,43) R(0) -> Twist prove / verify; the squares table 0..7 with
 * lookups [3, 5, 0, 7] -> Shout prove / verify.  Prints the canonical proof bytes (the values tests/golden/appendix_c.json pins).
 *
 *   make -C multilinear-map-cryptography_b200
 *   gcc -std=c99 -Iinclude examples/demo.c -o demo -Lmultilinear-map-cryptography_b200 -ltsgpu -Wl,-rpath,$PWD/multilinear-map-cryptography_b200 && ./demo
 *
 * Needs a CUDA device: tsgpu_init fails without one (there is no CPU fallback). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "tsgpu.h"

static void die(tsgpu_ctx* ctx, const char* what, int rc) {
    fprintf(stderr, "%s failed (%d): %s\n", what, rc, ctx ? tsgpu_last_error(ctx) : "no context");
    exit(1);
}

static void print_proof(const char* name, const tsgpu_proof* p) {
    size_t n = tsgpu_proof_bytes(p, NULL, 0);
    unsigned char* b = (unsigned char*)malloc(n);
    tsgpu_proof_bytes(p, b, n);
    printf("%s proof: %zu bytes, %zu sum-check rounds, %zu openings\n  ", name, n, tsgpu_proof_num_rounds(p), tsgpu_proof_num_openings(p));
    for (size_t i = 0; i < 64 && i < n; ++i) printf("%02x", b[i]);
    printf("...\n");
    free(b);
}

int main(void) {
    tsgpu_ctx* ctx = NULL;
    int rc = tsgpu_init(0, NULL, &ctx);
    if (rc) die(NULL, "tsgpu_init", rc);

    tsgpu_params* params = NULL;                                   /* setup_params(3): 8 cells, max 32 operations */
    if ((rc = tsgpu_setup_params(ctx, 3, &params))) die(ctx, "setup_params", rc);
    printf("setup_params(3): max_operations = %zu\n", tsgpu_params_max_operations(params));

    /* MemoryTrace::new(8): write(0, 42), write(1, 100), read(0), read(1), write(0, 43), read(0) - reads return the memory content */
    const uint64_t addresses[6] = {0, 1, 0, 1, 0, 0};
    const uint64_t raw_values[6] = {42, 100, 42, 100, 43, 43};
    const uint8_t is_write[6] = {1, 1, 0, 0, 1, 0};
    tsgpu_fr values[6];
    tsgpu_fr_from_u64(raw_values, 6, values);                      /* FieldElement::from(u64) */
    tsgpu_proof* twist = NULL;
    if ((rc = tsgpu_twist_prove(ctx, params, addresses, values, is_write, 6, &twist))) die(ctx, "Twist::prove", rc);
    int ok = 0;
    if ((rc = tsgpu_twist_verify(ctx, params, twist, &ok))) die(ctx, "Twist::verify", rc);
    print_proof("Twist", twist);
    printf("Twist::verify -> %s\n", ok ? "true" : "false");

    /* LookupTable of squares 0..7, lookups 3, 5, 0, 7 */
    uint64_t squares[8];
    for (int i = 0; i < 8; ++i) squares[i] = (uint64_t)(i * i);
    tsgpu_fr entries[8];
    tsgpu_fr_from_u64(squares, 8, entries);
    const uint64_t lookups[4] = {3, 5, 0, 7};
    tsgpu_proof* shout = NULL;
    if ((rc = tsgpu_shout_prove(ctx, params, entries, 8, lookups, 4, &shout))) die(ctx, "Shout::prove", rc);
    int ok2 = 0;
    if ((rc = tsgpu_shout_verify(ctx, params, shout, &ok2))) die(ctx, "Shout::verify", rc);
    print_proof("Shout", shout);
    printf("Shout::verify -> %s\n", ok2 ? "true" : "false");

    /* the reference's limit: Err(InvalidParameters("Too many operations")) beyond max_operations (twist.rs:108-112) */
    {
        uint64_t many_a[40] = {0}; tsgpu_fr many_v[40]; uint8_t many_w[40];
        memset(many_v, 0, sizeof many_v); memset(many_w, 1, sizeof many_w);
        tsgpu_proof* none = NULL;
        rc = tsgpu_twist_prove(ctx, params, many_a, many_v, many_w, 40, &none);
        printf("40 operations with max 32 -> error %d: %s\n", rc, tsgpu_last_error(ctx));
    }
    printf("%llu kernel launches\n", (unsigned long long)tsgpu_launch_count(ctx));
    tsgpu_proof_free(twist); tsgpu_proof_free(shout);
    tsgpu_params_free(ctx, params);
    tsgpu_destroy(ctx);
    return ok && ok2 ? 0 : 1;
}
