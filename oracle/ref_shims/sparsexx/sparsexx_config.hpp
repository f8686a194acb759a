// generated-config stand-in (MPI off) -- oracle/_ref glue
