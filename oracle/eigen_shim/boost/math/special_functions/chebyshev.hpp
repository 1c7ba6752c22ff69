// unused by the reference (included at include/utilities.h:13 only)
