/* Test double for the one symbol the XLA runtime exports to legacy custom calls:
 *   void XlaCustomCallStatusSetFailure(XlaCustomCallStatus*, const char* message, size_t message_len);
 * bl_xla_eval (biolith_b200/csrc/xla.cu) resolves it with dlsym(RTLD_DEFAULT, ...) -- loading this object with
 * RTLD_GLOBAL makes the failure path observable without jaxlib. */
#include <stddef.h>
#include <string.h>

static char g_message[512];
static void* g_status;

void XlaCustomCallStatusSetFailure(void* status, const char* message, size_t message_len) {
  size_t n = message_len < sizeof(g_message) - 1 ? message_len : sizeof(g_message) - 1;
  memcpy(g_message, message, n);
  g_message[n] = 0;
  g_status = status;
}

const char* bl_test_status_message(void) { return g_message[0] ? g_message : NULL; }
void* bl_test_status_token(void) { return g_status; }
void bl_test_status_reset(void) { g_message[0] = 0; g_status = NULL; }
