/* Embeds yart_b200/data/tables.bin (LUT values + Sobol dims 0-1) into the library. */
    .section .rodata
    .global yb_tables_start
    .global yb_tables_end
    .balign 16
yb_tables_start:
    .incbin YB_TABLES_PATH
yb_tables_end:
    .section .note.GNU-stack,"",@progbits
