// forwards to the single-header DIY stand-in (test infrastructure only)
#include <diy/stub_all.hpp>
