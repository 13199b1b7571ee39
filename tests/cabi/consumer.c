/* A plain C99 consumer of include/gcmf.h: proves that the drop-in boundary is usable without C++, Python or torch.
 * Built and run by tests/test_host_api.py (no GPU needed: only argument checks and version queries are made). */
#include <stdio.h>
#include <string.h>

#include "gcmf.h"

int main(void) {
    gcmf_plan_desc desc;
    gcmf_plan* plan = NULL;
    size_t bytes = 0;
    int rc;

    if (gcmf_version() != 1) return 10;
    printf("arch=%d\n", gcmf_sm_arch());

    memset(&desc, 0, sizeof desc);
    desc.op = 99; /* no such operator family */
    desc.dtype = GCMF_F64;
    desc.ny = 8;
    desc.nx = 8;
    rc = gcmf_plan_create(&desc, &plan);
    if (rc == GCMF_OK || plan != NULL) return 11;
    printf("rc=%d msg=%s\n", rc, gcmf_last_error());
    if (strstr(gcmf_last_error(), "unknown op") == NULL) return 12;

    rc = gcmf_plan_create(NULL, &plan);
    if (rc == GCMF_OK) return 13;
    rc = gcmf_workspace_bytes(NULL, 1, &bytes);
    if (rc == GCMF_OK) return 14;
    gcmf_plan_destroy(NULL); /* must be a no-op */
    printf("ok\n");
    return 0;
}
