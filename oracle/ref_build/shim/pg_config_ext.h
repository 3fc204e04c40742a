#define PG_INT64_TYPE long int
