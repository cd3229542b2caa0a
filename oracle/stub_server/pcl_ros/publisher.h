// stand-in: see haf_ref_stubs.hpp (oracle/stub_server) -- TEST INFRASTRUCTURE ONLY
#include "haf_ref_stubs.hpp"
