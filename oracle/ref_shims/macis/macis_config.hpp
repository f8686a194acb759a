// generated-config stand-in (MPI/Boost off) -- oracle/_ref glue
