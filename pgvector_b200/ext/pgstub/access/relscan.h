/* pgstub: syntax-check stand-in for the PostgreSQL header of the same name (NOT PostgreSQL code) */
#include "postgres.h"
