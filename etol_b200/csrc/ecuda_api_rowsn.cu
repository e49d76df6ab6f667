// ecuda_api_rowsn.cu -- second translation unit of ecuda_api.cu: the launchers (and with them the instantiations) of
// the N-specialised kernel family, compiled in parallel with the rest (see the note at the top of ecuda_api.cu).
#define ECUDA_TU_ROWSN 1
#include "ecuda_api.cu"
