#include "grb_host.h"
#include <cstdio>
extern "C" int grb_parse_card(const char *, const char *, int, const char *, grb_process_desc *, char *errbuf, int n) {
  if (errbuf && n > 0) snprintf(errbuf, n, "card parser not built yet");
  return GRB_ERR_UNSUPPORTED;
}
